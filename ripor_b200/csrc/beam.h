// Beam state shared by beam.cu (the beam-step / finalize kernels) and engine.cu (the decoder reads the
// KV ancestry table the beam step maintains).
#pragma once
#include "rb_common.h"
#include "trie.h"

#define RB_TAIL_MAX_L 32   // the forced-tail attention kernels hold a lineage's <= 32 key positions in one tile

struct rb200_beam {
  int device = 0, max_batch = 0, max_nb = 0, nb = 0, L = 0, V = 0;
  int batch = 0;   // queries of the batch in flight
  int step = 0;    // number of steps taken since reset
  int cur = 0;     // which half of the double buffers holds the current state
  double* scores[2] = {nullptr, nullptr};       // [R] float64 running beam scores (generation.py:463,505)
  rb::TrieState* state[2] = {nullptr, nullptr};  // [R] trie state of every beam
  int32_t* hist[2] = {nullptr, nullptr};        // [R, L] tokens chosen so far (input_ids[:, 1:], generation.py:511)
  int32_t* anc[2] = {nullptr, nullptr};         // [R, L] anc[r][p] = row whose K/V at position p belongs to r's lineage
  int32_t* parent = nullptr;                    // [R] in-query beam index chosen at the last step
  int32_t* token = nullptr;                     // [R] token chosen at the last step
  // ---- per-query forced tail (see below) ----
  int32_t* qstate = nullptr;                    // [max_batch] 0: still stepping; t > 0: frozen after t steps
  int32_t* qlist = nullptr;                     // [max_batch] the queries still stepping, ascending (compact -> query)
  int32_t* fz_list = nullptr;                   // [max_batch] frozen queries in freeze order
  int32_t* counts = nullptr;                    // [2] {queries still stepping, queries frozen}
  double* fz_scores = nullptr;                  // state of the frozen queries' beams (same indexing as the halves)
  rb::TrieState* fz_state = nullptr;
  int32_t* fz_hist = nullptr;
  int32_t* fz_anc = nullptr;
  int n_active = 0, n_frozen = 0;               // host mirror of counts (valid after beam_compact + read-back)
  bool compacted = false;                       // qlist differs from the identity
};

// ---- forced tail (engine.cu) --------------------------------------------------------------------------------------
// Once every beam of a QUERY sits on a single trie leaf, the rest of its DocIDs is determined by the trie: each later
// step keeps exactly its nb beams (their one valid child each), only their scores and their order change. Such a query
// is frozen: it leaves the step loop, and after the loop the remaining positions of all frozen queries run as ONE
// teacher-forced pass. Rows of that pass are position-major and ragged: block p holds the beams of the queries frozen
// at or before p, in freeze order, so row(frozen row r', position p) = off[p] + r'.
namespace rb {

struct TailLayout {
  int off[RB_TAIL_MAX_L + 1];   // off[p] = first row of position block p; off[P] = total rows
  int P;                        // positions of a finished lineage (= max_new_tokens)
};

// one beam step over the queries still stepping (compact row order = qlist order). allow_freeze: queries whose beams
// all sit on a single leaf after this step are frozen (state -> fz_*). next_x rows are written at ORIGINAL row ids.
int beam_step(rb200_beam* bm, const rb200_trie* trie, const float* logits, int rows_per_query, int apply_log_softmax,
              const float* embed_table, float* next_x, int d_model, int allow_freeze, cudaStream_t s);
// rebuild qlist / fz_list / counts from qstate after a step that may have frozen queries
int beam_compact(rb200_beam* bm, cudaStream_t s);
// x[m, :] = x_full[orig(m), :] for the compact rows of the queries still stepping
int launch_gather_rows(const rb200_beam* bm, const float* x_full, float* x, int d_model, cudaStream_t s);
// decoder inputs of every (frozen row, position >= its freeze step) and the forced tokens into fz_hist
int launch_tail_prepare(rb200_beam* bm, const rb200_trie* trie, const TailLayout& lay, const float* const* in_tabs_dev,
                        const float* start_emb, float* x, int d_model, cudaStream_t s);
// replays the remaining steps of every frozen query on the logits of the pass: float64 score adds in step order and
// the per-step re-ranking of the beams, exactly as the step loop would do them (generation.py:463-507)
int launch_tail_finish(rb200_beam* bm, const rb200_trie* trie, const TailLayout& lay, const float* logits,
                       int apply_log_softmax, cudaStream_t s);
// teacher-forced scoring (rb200_engine_forward): every query "frozen" at step 0 with the given tokens as history
int beam_force_tokens(rb200_beam* bm, int batch, int nb, int T, const int32_t* tokens_dev, cudaStream_t s);
// scores[r] = sum_p logits[row(r, p), tokens[r][p]] (float64 accumulate, fp32 result)
int launch_forced_scores(const rb200_beam* bm, const TailLayout& lay, const float* logits, float* scores,
                         cudaStream_t s);
}  // namespace rb
