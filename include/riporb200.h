/*
 * riporb200.h — C ABI of the B200-native constrained-beam-search retrieval path.
 *
 * This is the drop-in boundary for ONE path of HansiZeng/RIPOR: query tokens -> T5 encoder ->
 * L-step prefix-constrained beam search over the DocID trie -> ranked DocID list
 * (reference t5_pretrainer/evaluate.py:396-487, t5_pretrainer/tasks/generation.py:35-677,
 * t5_pretrainer/modeling/t5_generative_retriever.py:194-262,295-450).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++ or torch types.
 *   - every function returns 0 on success or a negative rb200_status; rb200_last_error() gives the
 *     message of the last failure on the calling thread. Nothing throws across the boundary.
 *   - "_dev" pointers are device pointers owned by the caller (PyTorch allocates them); "_host"
 *     pointers are host memory owned by the caller. Handles own only their internal tables/workspaces.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream).
 *   - one handle per host thread / stream; handles are not re-entrant.
 */
#ifndef RIPORB200_H
#define RIPORB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  RB200_OK = 0,
  RB200_ERR_INVALID = -1,   /* bad argument (the reference raises ValueError / assert) */
  RB200_ERR_CUDA = -2,      /* CUDA runtime/driver failure; message holds cudaGetErrorString */
  RB200_ERR_IO = -3,        /* file could not be read/written */
  RB200_ERR_STATE = -4,     /* call order violated (e.g. trie not uploaded, weights not packed) */
  RB200_ERR_NOMEM = -5
} rb200_status;

const char* rb200_version(void);
const char* rb200_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * DocID trie. Replaces the reference's per-level {"-1_c1_.._ci": [next ids]} dicts
 * (evaluate.py:411-424, aq_preprocess/build_list_smtid_to_nextids.py:23-36), the scipy-CSR mask tables of
 * PrefixConstrainLogitProcessorFastSparse.__init__ (generation.py:604-642) and the smtid -> [docids]
 * dict (evaluate.py:439-446) by one flattened structure over the lexicographically sorted unique codes.
 * ------------------------------------------------------------------------------------------------ */
typedef struct rb200_trie rb200_trie;

typedef struct {
  int64_t n_docs;        /* rows given to rb200_trie_build */
  int64_t n_unique;      /* distinct code rows (= distinct smtids = leaves) */
  int64_t n_nodes;       /* explicit (bitmap) nodes: prefixes whose range holds > 32 unique codes */
  int64_t n_children;    /* child slots of the explicit nodes */
  int64_t bytes;         /* bytes of all tables (what rb200_trie_upload places in HBM) */
  int32_t L;             /* DocID length */
  int32_t V;             /* codebook size */
  int32_t code_bytes;    /* 1 (V <= 256) or 2 */
  int32_t on_device;     /* device ordinal the tables were uploaded to, -1 if host only */
} rb200_trie_info;

/* codes_host: row-major [n_docs, L] of uint8 (code_bytes=1) or uint16 (code_bytes=2); row i is the
 * i-th entry of docid_to_smtid.json without its leading -1. All values must be < V. Host only. */
int rb200_trie_build(const void* codes_host, int code_bytes, int64_t n_docs, int L, int V, int n_threads,
                     rb200_trie** out);
int rb200_trie_free(rb200_trie* trie);
int rb200_trie_get_info(const rb200_trie* trie, rb200_trie_info* info);
/* counts_host[i] = number of distinct prefixes of length i (i = 0..L-1): the figure the reference prints
 * as "{i}-th step has {n} effective smtid" (evaluate.py:425-426). */
int rb200_trie_level_counts(const rb200_trie* trie, int64_t* counts_host);
/* versioned binary cache replacing list_smtid_to_nextids.pkl (evaluate.py:404-432). */
int rb200_trie_save(const rb200_trie* trie, const char* path);
int rb200_trie_load(const char* path, rb200_trie** out);
/* same, with a caller-chosen 64-bit tag of the SOURCE the trie was built from (e.g. size/mtime/hash of
 * docid_to_smtid.json) stored in the header, so that a stale cache is detected instead of silently mapping beams to
 * the wrong documents. The file is written under a temporary name and renamed into place. */
int rb200_trie_save_tagged(const rb200_trie* trie, const char* path, uint64_t source_tag);
int rb200_trie_load_tagged(const char* path, uint64_t* source_tag, rb200_trie** out);
/* Host walk with the semantics of PrefixConstrainLogitProcessorFastSparse.__call__ (generation.py:666-677):
 * input_ids_host [R, T] int64 (column 0 is the decoder start token and is ignored), mask_host [R, V]
 * float64: 1.0 on allowed next tokens, all-zero row for a prefix that is not in the trie. */
int rb200_trie_mask_host(const rb200_trie* trie, const int64_t* input_ids_host, int64_t R, int T,
                         double* mask_host);
/* leaf (= index into the sorted unique codes; rb200_beam_finalize returns a [lo, hi) range of leaves per
 * output row) -> the rows of codes_host that carry that smtid, in input order (evaluate.py:439-446 keeps
 * json order). */
int rb200_trie_leaf_docs(const rb200_trie* trie, int64_t leaf, const int64_t** docs_host, int64_t* n);
/* exact-match lookup of a full code row; *leaf = -1 if absent. code_host has L entries (int32). */
int rb200_trie_find_leaf(const rb200_trie* trie, const int32_t* code_host, int64_t* leaf);
/* copy the tables to HBM of `device` (idempotent per device; a handle may be uploaded to several devices - the
 * reference runs one process per GPU, evaluate.py:463-470, a single-process caller uploads once per device). The
 * device entry points below use the copy on the calling thread's current device (mask, leaf expansion) or on the
 * beam state's / engine's device. */
int rb200_trie_upload(rb200_trie* trie, int device);
/* Device form of the smtid -> docids mapping (evaluate.py:118-128, 439-446): leaf_ranges_dev int32 [n, 2] as
 * returned by rb200_beam_finalize / rb200_engine_search; docs_dev int64 [n, max_docs_per_row] receives the input rows
 * (= positions in docid_to_smtid.json) of the documents under each range in json order, padded with -1;
 * counts_dev int32 [n] the true number of documents (0: the row is not in the trie, the reference prints and skips
 * it; > max_docs_per_row: the row was not expanded, use rb200_trie_leaf_docs on the host). The leaf table
 * (8 bytes per document) is uploaded on first use. */
int rb200_trie_leaf_expand(rb200_trie* trie, const int32_t* leaf_ranges_dev, int64_t n, int max_docs_per_row,
                           int64_t* docs_dev, int32_t* counts_dev, void* stream);
/* Device form of the mask processor call: same semantics as rb200_trie_mask_host, all pointers on device. */
int rb200_trie_mask_device(const rb200_trie* trie, const int64_t* input_ids_dev, int64_t R, int T,
                           double* mask_dev, void* stream);

/* ------------------------------------------------------------------------------------------------
 * On-disk inputs. docid_to_smtid.json = {"<docid>": [-1, c1, .., cL], ...} (written by the reference's
 * aq_preprocess/create_customized_smtid_file.py:47-58, read with ujson.load at evaluate.py:400-401): a streaming
 * parser into a codes matrix and a docid string table in file order. max_len > 0 keeps only the first max_len codes
 * of every document (evaluate.py:443: smtids[1:1+max_new_token_for_docid]).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rb200_docid_table rb200_docid_table;
int rb200_docid_json_open(const char* path, int max_len, rb200_docid_table** out);
int rb200_docid_json_free(rb200_docid_table* table);
int rb200_docid_json_info(const rb200_docid_table* table, int64_t* n_docs, int32_t* L, int32_t* max_code,
                          int64_t* key_bytes);
/* int32 [n_docs, L] codes (valid until rb200_docid_json_free). */
int rb200_docid_json_codes(const rb200_docid_table* table, const int32_t** codes_host);
/* docid i = key_bytes[key_offsets[i] .. key_offsets[i+1]) (UTF-8, not terminated). */
int rb200_docid_json_keys(const rb200_docid_table* table, const char** key_bytes_host, const int64_t** key_offsets_host);
int rb200_trie_build_from_table(const rb200_docid_table* table, int V, int n_threads, rb200_trie** out);
/* Residual-quantiser codes as faiss packs them: M codes of `bits` bits each, LSB first, in code_size bytes per
 * vector; the reference unpacks them with faiss.BitstringReader when bits != 8
 * (aq_preprocess/create_customized_smtid_file.py:38-45). out_host int32 [n, M]. */
int rb200_unpack_codes(const uint8_t* packed_host, int64_t n, int64_t code_size, int M, int bits, int32_t* out_host);

/* ------------------------------------------------------------------------------------------------
 * Beam state + beam step. Replaces, per decoding step, generation.py:453-463 (optional log-softmax,
 * mask add in float64, beam-score add), :485-492 (top-2*nb, div/mod), HF BeamSearchScorer.process
 * (:496-507, "first nb of the 2*nb"), :511 (input_ids reorder) and the per-step host round trip of the
 * mask processor (:666-677); rb200_beam_finalize replaces BeamSearchScorer.finalize (:532-540).
 * ------------------------------------------------------------------------------------------------ */
typedef struct rb200_beam rb200_beam;

int rb200_beam_create(int device, int max_batch, int num_beams, int L, int V, rb200_beam** out);
int rb200_beam_free(rb200_beam* beam);
/* start a batch of `batch` queries: beam 0 score 0, beams 1.. score -1e9 (generation.py:418-420),
 * every beam at the trie root, step counter 0. */
int rb200_beam_reset(rb200_beam* beam, const rb200_trie* trie, int batch, void* stream);
/* same with fewer beams than the state was created for (1 <= num_beams <= the num_beams of rb200_beam_create): the
 * reference builds a new BeamSearchScorer per generate call (generation.py:222-229), so the beam width is a per-call
 * argument there too. */
int rb200_beam_reset_beams(rb200_beam* beam, const rb200_trie* trie, int batch, int num_beams, void* stream);
/* One step. logits_dev: fp32 [batch*rows_per_query, V]; rows_per_query is 1 when all beams of a query
 * share one logits row (step 0: identical prefixes) or num_beams. apply_log_softmax mirrors
 * apply_log_softmax_for_scores. If next_embed_table_dev != NULL (fp32 [V, d_model], the reference's
 * list_decoder_embeds[t].weight, t = the step just taken) the kernel also gathers the next decoder input
 * rows into next_x_dev fp32 [batch*num_beams, d_model] (t5_generative_retriever.py:209-211). */
int rb200_beam_step(rb200_beam* beam, const rb200_trie* trie, const float* logits_dev, int rows_per_query,
                    int apply_log_softmax, const float* next_embed_table_dev, float* next_x_dev, int d_model,
                    void* stream);
/* sequences_dev int64 [batch*num_return, L+1] (col 0 = decoder start id 0), scores_dev fp32
 * [batch*num_return] = float32(beam_score / (L+1)^length_penalty), rows ordered like HF finalize
 * (descending score, ties in reverse beam order); leaf_dev int32 [batch*num_return, 2] = the [lo, hi)
 * range of trie leaves below the row's code prefix (one leaf when the prefix is a full DocID; lo == hi
 * when the row is not in the trie: the reference drops those at evaluate.py:121-122).
 * num_return <= num_beams else RB200_ERR_INVALID (generation.py:216-217). */
int rb200_beam_finalize(rb200_beam* beam, const rb200_trie* trie, int num_return, double length_penalty,
                        int64_t* sequences_dev, float* scores_dev, int32_t* leaf_dev, void* stream);
/* read-only views of the current state (device pointers, valid until the next step/reset):
 * what = 0: beam_scores f64 [batch*nb]; 1: parent (in-query beam index) i32 [batch*nb];
 * 2: last tokens i32 [batch*nb]; 3: token history i32 [batch*nb, L]; 4: KV ancestry i32 [batch*nb, L];
 * 5: trie state i32 [batch*nb, 4] (lo, hi, node, 0). */
/* Forced-tail score replay (the part of the engine's forced tail that is beam arithmetic), exposed for parity tests:
 * the state must sit at step t with every beam of every query on a single trie leaf. tail_logits_dev: fp32
 * [T, batch*num_beams, V], block j = the logits of step t+j for the beams IN THEIR ORDER AT STEP t (a lineage's
 * logits do not depend on the slot it occupies). Leaves scores, token history, trie state and beam order exactly
 * where T more rb200_beam_step calls on the same logits would (float64 adds in step order and the per-step
 * re-ranking of generation.py:463-507). t + T <= 32. */
int rb200_beam_forced_tail(rb200_beam* beam, const rb200_trie* trie, int T, const float* tail_logits_dev,
                           int apply_log_softmax, void* stream);
int rb200_beam_view(const rb200_beam* beam, int what, const void** ptr_dev);
int rb200_beam_current_step(const rb200_beam* beam);

/* ------------------------------------------------------------------------------------------------
 * Engine: the T5 encoder + KV-cached decoder + beam search for one GPU. Replaces
 * generate_for_constrained_prefix_beam_search (generation.py:35-251) and the model forward it drives
 * (t5_generative_retriever.py:295-450) for inference.
 * ------------------------------------------------------------------------------------------------ */
typedef struct rb200_engine rb200_engine;

typedef enum {
  RB200_PREC_FP32 = 0,    /* fp32 FFMA GEMMs (no tensor cores); exact-arithmetic reference mode */
  RB200_PREC_TF32X3 = 1,  /* tcgen05 kind::tf32, error-compensated 3-MMA split: fp32-grade, parity mode */
  RB200_PREC_BF16X3 = 2,  /* tcgen05 kind::f16 (bf16), 3-MMA split: ~2^-16 relative */
  RB200_PREC_TF32 = 3,    /* single tf32 MMA */
  RB200_PREC_BF16 = 4,    /* single bf16 MMA: throughput mode, not parity-safe */
  RB200_PREC_FP16X3 = 5   /* tcgen05 kind::f16 (fp16), 3-MMA split: tf32x3's 11-bit planes at half the bytes and
                             twice the MMA rate. fp16 range: activations beyond +-65000 raise an overflow flag and
                             the search returns NaN scores (use tf32x3 for such checkpoints) */
} rb200_precision;

typedef struct {
  int32_t d_model, num_heads, d_kv, d_ff;
  int32_t num_layers, num_decoder_layers;
  int32_t vocab_size;                 /* encoder token vocabulary (shared.weight rows) */
  int32_t num_buckets, max_distance;  /* relative attention */
  float layer_norm_eps;
  int32_t decoder_vocab_size;         /* V; the reference requires it uniform (evaluate.py:433-436) */
  int32_t docid_len;                  /* number of codebook positions the model has tables for */
  int32_t shared_output_input_embeds; /* t5_generative_retriever.py:254-258 */
  int32_t scaleup_output_hidden;      /* t5_generative_retriever.py:427-428 */
  int32_t max_batch;                  /* capacities of the workspaces (see rb200_engine_resize): queries per call, */
  int32_t max_beams;                  /* beams per query, */
  int32_t max_src_len;                /* padded query length S */
  int32_t precision;                  /* rb200_precision */
  int32_t device;
} rb200_engine_config;

int rb200_engine_create(const rb200_engine_config* cfg, rb200_engine** out);
int rb200_engine_free(rb200_engine* eng);
/* Re-allocate the shape-dependent workspaces (activations, KV cache, beam state) for new capacities; the packed
 * weights stay. Every call below accepts any batch <= max_batch, num_beams <= max_beams, S <= max_src_len. */
int rb200_engine_resize(rb200_engine* eng, int max_batch, int max_beams, int max_src_len);
/* Hand one fp32 tensor of the HF state dict to the engine by its state-dict key,
 * e.g. "decoder.block.3.layer.1.EncDecAttention.q.weight", "list_output_embeds.7.weight",
 * "start_token_embed". data_dev is a device pointer to contiguous fp32 in the HF layout; the engine
 * copies/packs it (fused QKV, precision split) so the caller may free it afterwards. */
int rb200_engine_set_weight(rb200_engine* eng, const char* name, const float* data_dev, int64_t numel,
                            void* stream);
/* finish packing; fails with RB200_ERR_STATE listing the first missing tensor. */
int rb200_engine_finalize_weights(rb200_engine* eng, void* stream);
int64_t rb200_engine_workspace_bytes(const rb200_engine* eng);

/* Whole path on device buffers. input_ids_dev / attention_mask_dev: int64 [batch, S] (the reference
 * passes .long() tensors, evaluate.py:105-106). Outputs as rb200_beam_finalize. */
int rb200_engine_search(rb200_engine* eng, const rb200_trie* trie, const int64_t* input_ids_dev,
                        const int64_t* attention_mask_dev, int batch, int S, int num_beams, int max_new_tokens,
                        int num_return, int apply_log_softmax, int64_t* sequences_dev, float* scores_dev,
                        int32_t* leaf_dev, void* stream);
/* Same with HOST buffers (pinned or pageable): copies inputs H2D, runs, copies results D2H and
 * synchronises the stream before returning. This is the end-to-end call bench.py times as `e2e`. */
int rb200_engine_search_host(rb200_engine* eng, const rb200_trie* trie, const int64_t* input_ids_host,
                             const int64_t* attention_mask_host, int batch, int S, int num_beams,
                             int max_new_tokens, int num_return, int apply_log_softmax,
                             int64_t* sequences_host, float* scores_host, int32_t* leaf_host, void* stream);

/* Pieces of the path, exposed for parity tests and for callers that keep their own loop. */
/* encoder + cross-attention K/V projection for a batch; must precede rb200_engine_decode_step. */
int rb200_engine_encode(rb200_engine* eng, const int64_t* input_ids_dev, const int64_t* attention_mask_dev,
                        int batch, int S, int num_beams, void* stream);
/* encoder last_hidden_state fp32 [batch, S, d_model] of the last rb200_engine_encode (device pointer). */
int rb200_engine_encoder_states(const rb200_engine* eng, const float** states_dev);
/* decoder position t for all rows; rows = batch (t == 0) or batch*num_beams. `beam` supplies the KV
 * ancestry and the decoder inputs were placed by the previous rb200_beam_step (or the start embedding).
 * logits_dev fp32 [rows, V]. */
int rb200_engine_decode_step(rb200_engine* eng, const rb200_beam* beam, int t, float* logits_dev, void* stream);
/* the beam state the engine owns (created for max_batch x max_beams). */
int rb200_engine_beam(rb200_engine* eng, rb200_beam** beam);
/* fp32 [max_batch*max_beams, d_model] buffer rb200_engine_decode_step takes its inputs from for t >= 1: pass it as
 * next_x_dev to rb200_beam_step (rows at beam-row ids). */
int rb200_engine_next_input(const rb200_engine* eng, float** next_x_dev);
/* Teacher-forced decoder pass over GIVEN DocID tokens, no beam search: the model forward the reference runs when it
 * is handed decoder_input_ids (t5_generative_retriever.py:295-450) and the sum rerank_forward takes over it
 * (:794-798). tokens_dev int32 [batch*rows_per_query, T]: tokens[r][p] is the code at position p of row r's DocID;
 * position p's decoder input is the start embedding (p = 0) or list_decoder_embeds[p-1][tokens[r][p-1]]. Outputs, each
 * optional (NULL), all POSITION-MAJOR with R = batch*rows_per_query rows per position:
 *   logits_dev fp32 [T, R, V]        = the reference's `logits` list (position p with the p-th output table),
 *   hidden_dev fp32 [T, R, d_model]  = decoder_last_hidden_state (after the final layer norm and the optional d^-1/2),
 *   scores_dev fp32 [R]              = sum_p logits[p][r][tokens[r][p]]   (= rerank_forward's score).
 * The rows of a query share its encoder pass. T <= 32. */
int rb200_engine_forward(rb200_engine* eng, const int64_t* input_ids_dev, const int64_t* attention_mask_dev, int batch,
                         int S, int rows_per_query, const int32_t* tokens_dev, int T, float* logits_dev,
                         float* hidden_dev, float* scores_dev, void* stream);
/* number of kernel launches issued by the last rb200_engine_search* call. */
int64_t rb200_engine_last_launch_count(const rb200_engine* eng);
/* Forced tail: once every beam of a QUERY sits on a single trie leaf, the rest of its DocIDs is determined; the query
 * leaves the step loop and the remaining positions of all such queries are evaluated in one teacher-forced pass after
 * the loop (same per-row arithmetic as the steps of generation.py:423-530 they replace, far fewer and larger
 * kernels). Returns the first step at which a query of the last search was frozen, or -1 if every query ran step by
 * step to the end (RB200_TAIL=0 disables freezing). */
int rb200_engine_last_tail_step(const rb200_engine* eng);
/* frozen_at_host[t] (t < n) = queries of the last search frozen after t steps; *tail_rows_host = rows of its
 * forced-tail pass (sum over frozen queries of num_beams * remaining positions). */
int rb200_engine_last_freeze_histogram(const rb200_engine* eng, int32_t* frozen_at_host, int n,
                                       int64_t* tail_rows_host);
/* Measurement aid for bench.py's roofline leg: with profiling on, every GEMM launch of the engine is
 * bracketed by CUDA events on its stream. get_profile synchronises and returns the summed GEMM device time
 * (ms), the algorithmic FLOPs (2*M*N*K per launch) and the launch count since profiling was switched on. */
int rb200_engine_set_profiling(rb200_engine* eng, int on);
int rb200_engine_get_profile(rb200_engine* eng, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);
/* algorithmic HBM bytes of the same launches: operand planes read + result written (a residual add counts the old
 * value's read and the new value's write; both are done by L2's reduce-add). */
int rb200_engine_get_profile_bytes(const rb200_engine* eng, double* gemm_bytes);

/* One GEMM of the engine's family, for kernel-level parity tests and the roofline microbench:
 * C[M,N] = A[M,K] * W[N,K]^T (+ C if accumulate) in the given precision. A, W, C fp32 device, row-major. */
int rb200_gemm(int precision, const float* A_dev, const float* W_dev, float* C_dev, int64_t M, int64_t N,
               int64_t K, int accumulate, int relu, void* stream);

/* Roofline microbench of one GEMM shape: `iters` back-to-back launches (epilogue 0 store, 1 residual, 2 relu->planes)
 * over operands rotated through `rotate_mb` MiB of copies (so weights do not stay in L2 between launches, as in a
 * real decoder step), timed with CUDA events on `stream`. avg_us = mean device time per launch. */
int rb200_gemm_bench(int precision, int64_t M, int64_t N, int64_t K, int epilogue, int iters, int rotate_mb,
                     double* avg_us, void* stream);

/* HF T5Attention._relative_position_bucket for one relative position (key - query), in the float32
 * arithmetic torch uses; the engine builds its bias tables with it (exposed for parity tests). */
int rb200_relative_position_bucket(int relative_position, int bidirectional, int num_buckets, int max_distance);

#ifdef __cplusplus
}
#endif
#endif /* RIPORB200_H */
