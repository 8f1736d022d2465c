// Engine: T5 encoder + KV-cached decoder + constrained beam search for one GPU, behind the C ABI.
//
// Replaces generate_for_constrained_prefix_beam_search (reference t5_pretrainer/tasks/generation.py:35-251)
// and the model forward it drives every step (t5_pretrainer/modeling/t5_generative_retriever.py:295-450).
// Differences in HOW (not in results): the encoder states are never expanded x num_beams
// (generation.py:231-233) - cross K/V are projected once per query and shared by its beams; the decoder
// runs one position per step against a KV cache addressed through the beam ancestry table instead of
// re-running the whole prefix (finding 4) and instead of _reorder_cache copies
// (t5_generative_retriever.py:484-512); step 0 runs one row per query because all beams are identical.
#include <map>
#include <set>
#include <string>
#include <vector>

#include "beam.h"
#include "kernels.h"

using rb::ActOut;
using rb::GemmArgs;

namespace {

struct Packed {          // a weight matrix [N, K] in the planes of the engine precision
  void* ptr = nullptr;
  int64_t plane = 0;     // elements between planes
  int64_t N = 0, K = 0;
};

struct Layer {
  Packed qkv, o, wi, wo;           // self-attention fused q|k|v, output, feed-forward
  Packed cq, co;                   // decoder cross-attention query / output
  float* ln0 = nullptr;
  float* ln1 = nullptr;
  float* ln2 = nullptr;
};

}  // namespace

// The workspaces of one batch in flight. (Round 1 also had a second lane on a second stream for half of the batch;
// measured slower on B200 - the GEMMs are bound by operand delivery, not by idle SMs - and removed.)
struct Lane {
  int64_t Mcap = 0, Rcap = 0, BScap = 0;
  float* x = nullptr;              // [Mcap, d] residual stream
  float* x_full = nullptr;         // [Rcap, d] next decoder inputs of every beam row, at original row ids
  void* xn = nullptr;              // ActBuf [planes][Mcap][d]
  float* qkv = nullptr;            // [Mcap, 3*inner]
  float* q2 = nullptr;             // [Mcap, inner]
  void* ctx = nullptr;             // ActBuf [planes][Mcap][inner]
  void* hbuf = nullptr;            // ActBuf [planes][Mcap][dff]
  float* logits = nullptr;         // [Rcap, V]
  float* cross_kv = nullptr;       // [Nl][BScap][2*inner]: per decoder layer, K | V of every source position
  float* enc_out = nullptr;        // [BScap, d]
  float* cache_k = nullptr;        // [Nl][Lmodel][Rcap][inner]
  float* cache_v = nullptr;
  float* ss[2] = {nullptr, nullptr};   // NormFold: [Mcap, np] per-row partial sums of x^2, two norm points alive
  rb200_beam* beam = nullptr;
  // batch in flight
  int B = 0, S = 0, nb = 0;
  const int64_t* cur_mask = nullptr;
};

struct rb200_engine {
  rb200_engine_config cfg;
  int mode = 0, planes = 1, elem = 4;
  int d = 0, inner = 0, dff = 0, H = 0, V = 0, Lmodel = 0;
  std::vector<Layer> enc, dec;
  Packed ckv;                      // all decoder layers' cross K|V projections: [Nl*2*inner, d]
  std::vector<Packed> out_tab;     // per position output table [V, d]
  std::vector<float*> in_tab;      // per position input table fp32 [V, d]
  float* shared_emb = nullptr;     // [vocab, d]
  float* start_emb = nullptr;      // [d]
  float* enc_final_ln = nullptr;
  float* dec_final_ln = nullptr;
  std::vector<float> enc_rel_host, dec_rel_host;   // [num_buckets, H]
  float* dec_bias = nullptr;       // [H, Lmodel]
  float* enc_bias = nullptr;       // [H, 2*S-1]
  int enc_bias_S = 0;
  std::set<std::string> have;
  bool finalized = false;
  // NormFold (kernels.h): layer-norm weights folded into the packed matrices that follow them
  struct Stash { float* src; Packed* w; int64_t row0, rows; const float* ln; float scale; };
  bool fold = false;
  int np = 0;
  std::vector<Stash> stash;
  Lane lanes[1];
  int64_t* ids_dev = nullptr;      // host-call staging
  int64_t* mask_dev = nullptr;
  int64_t* seq_dev = nullptr;
  float* score_dev = nullptr;
  int32_t* leaf_dev = nullptr;
  int64_t ws_bytes = 0, ws_only_bytes = 0;
  int64_t launches = 0;
  // optional per-GEMM event timing (rb200_engine_set_profiling)
  bool profiling = false;
  std::vector<cudaEvent_t> events;
  size_t ev_used = 0;
  double prof_flops = 0.0, prof_bytes = 0.0;

  int* overflow = nullptr;         // device flags: [0] activation overflow of the batch in flight, [1] weights
  // forced tail (beam.h): once every beam sits on a single trie leaf, the remaining positions run as one pass
  bool tail = false;
  float** in_tab_dev = nullptr;    // [Lmodel] device copy of in_tab
  int32_t* flag_host = nullptr;    // pinned [2]: {queries still stepping, queries frozen} after a beam step
  int last_tail_from = -1;         // first step at which a query of the last search was frozen (-1: none)
  int frozen_at[RB_TAIL_MAX_L + 1] = {0};   // last search: queries frozen after t steps
  int64_t last_tail_rows = 0;      // rows of the last forced-tail pass
  int ws_batch = 0, ws_beams = 0, ws_src = 0;   // capacities of the current workspaces

  ActOut act(const Lane& l, void* base, int64_t row_len) const {
    return ActOut{base, l.Mcap * row_len, mode, overflow};
  }
};

namespace {

int dev_alloc(rb200_engine* e, void** p, int64_t bytes) {
  RB_CUDA(cudaMalloc(p, (size_t)(bytes > 0 ? bytes : 16)));
  e->ws_bytes += bytes;
  return 0;
}

int alloc_packed(rb200_engine* e, Packed* w, int64_t N, int64_t K) {
  w->N = N; w->K = K; w->plane = N * K;
  return dev_alloc(e, &w->ptr, (int64_t)e->planes * N * K * e->elem);
}

// pack `rows` x K fp32 rows into row offset `row0` of packed weight w
int pack_rows(rb200_engine* e, Packed* w, int64_t row0, const float* src, int64_t rows, cudaStream_t s,
              const float* ln = nullptr, float extra = 1.0f) {
  if (e->fold && ln != nullptr) {
    // the layer-norm vector may not have been set yet: keep a copy and fold when the weights are finalized
    float* copy = nullptr;
    RB_CUDA(cudaMalloc((void**)&copy, (size_t)rows * w->K * 4));
    RB_CUDA(cudaMemcpyAsync(copy, src, (size_t)rows * w->K * 4, cudaMemcpyDeviceToDevice, s));
    e->stash.push_back({copy, w, row0, rows, ln, extra});
    return 0;
  }
  char* dst = static_cast<char*>(w->ptr) + row0 * w->K * e->elem;
  const float scale = rb::prec_is_fp16(e->mode) ? rb::kFp16WeightScale : 1.0f;   // (extra only applies when folding)
  return rb::launch_pack_planes(src, dst, rows * w->K, w->plane, e->mode, scale, e->overflow + 1, s);
}

int copy_f32(float* dst, const float* src, int64_t n, cudaStream_t s) {
  RB_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
  return 0;
}

int gemm(rb200_engine* e, const Lane& l, const void* A, int64_t a_row_len, const Packed& w, float* C, int64_t ldc,
         ActOut act, int64_t M, int epi, cudaStream_t s, rb::NormFold nf = rb::NormFold{}) {
  GemmArgs g;
  g.nf = nf;
  g.mode = e->mode;
  g.A = A; g.a_plane = l.Mcap * a_row_len;
  g.W = w.ptr; g.w_plane = w.plane;
  g.C = C; g.ldc = ldc; g.act = act;
  g.M = M; g.N = w.N; g.K = w.K; g.epilogue = epi;
  g.out_scale = rb::prec_is_fp16(e->mode) ? 1.0f / rb::kFp16WeightScale : 1.0f;
  if (!e->profiling) return rb::launch_gemm(g, s);
  // profiling pass: bracket every GEMM launch with events on its own stream (bench.py roofline leg)
  if (e->ev_used + 2 > e->events.size()) {
    const size_t old = e->events.size();
    e->events.resize(old + 1024);
    for (size_t i = old; i < e->events.size(); ++i) RB_CUDA(cudaEventCreate(&e->events[i]));
  }
  RB_CUDA(cudaEventRecord(e->events[e->ev_used], s));
  const int st = rb::launch_gemm(g, s);
  RB_CUDA(cudaEventRecord(e->events[e->ev_used + 1], s));
  e->ev_used += 2;
  e->prof_flops += 2.0 * (double)M * (double)w.N * (double)w.K;
  {   // algorithmic HBM bytes of this launch: operand planes in, result out (the residual is read and written by L2)
    const double pe = (double)e->planes * e->elem, mn = (double)M * (double)w.N;
    double out = 4.0 * mn;
    if (epi == rb::EPI_RESIDUAL) out = 8.0 * mn;
    else if (epi == rb::EPI_RELU_ACT || epi == rb::EPI_PLANES) out = pe * mn;
    else if (epi == rb::EPI_RESID_NORM) out = 8.0 * mn + pe * mn;
    e->prof_bytes += pe * ((double)M * (double)w.K + (double)w.N * (double)w.K) + out;
  }
  return st;
}

// rows [row0, row0 + rows) of a packed weight as a weight of its own (same planes, same plane distance)
Packed sub_rows(const rb200_engine* e, const Packed& w, int64_t row0, int64_t rows) {
  Packed p = w;
  p.ptr = static_cast<char*>(w.ptr) + row0 * w.K * e->elem;
  p.N = rows;
  return p;
}

// fp16x3: an activation left the fp16 range somewhere in this search -> make the result unmistakably invalid
__global__ void poison_on_overflow_kernel(float* scores, int n, const int* flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && *flag) scores[i] = __int_as_float(0x7fc00000);
}

int build_bias_tables(rb200_engine* e, int S, cudaStream_t s) {
  const auto& c = e->cfg;
  if (e->dec_bias == nullptr) {
    std::vector<float> t((size_t)e->H * e->Lmodel);
    for (int h = 0; h < e->H; ++h)
      for (int dist = 0; dist < e->Lmodel; ++dist)
        t[(size_t)h * e->Lmodel + dist] =
            e->dec_rel_host[(size_t)rb::relative_bucket(-dist, false, c.num_buckets, c.max_distance) * e->H + h];
    RB_TRY(dev_alloc(e, (void**)&e->dec_bias, (int64_t)t.size() * 4));
    RB_CUDA(cudaMemcpyAsync(e->dec_bias, t.data(), t.size() * 4, cudaMemcpyHostToDevice, s));
    RB_CUDA(cudaStreamSynchronize(s));
  }
  if (S > 0 && e->enc_bias_S != c.max_src_len) {
    // indexed by (key - query) + Smax - 1 for the workspace's max_src_len: valid for every batch's padded length
    const int Sm = c.max_src_len;
    std::vector<float> t((size_t)e->H * (2 * Sm - 1));
    for (int h = 0; h < e->H; ++h)
      for (int off = 0; off < 2 * Sm - 1; ++off)
        t[(size_t)h * (2 * Sm - 1) + off] =
            e->enc_rel_host[(size_t)rb::relative_bucket(off - (Sm - 1), true, c.num_buckets, c.max_distance) * e->H + h];
    if (e->enc_bias == nullptr) RB_CUDA(cudaMalloc((void**)&e->enc_bias, (size_t)e->H * (2 * Sm - 1) * 4));
    RB_CUDA(cudaMemcpyAsync(e->enc_bias, t.data(), t.size() * 4, cudaMemcpyHostToDevice, s));
    RB_CUDA(cudaStreamSynchronize(s));
    e->enc_bias_S = Sm;
  }
  return 0;
}

}  // namespace

namespace rb { void set_gemm_trace(unsigned long long* dev); }   // gemm_sm100_2cta.cu

static float half_bits_to_float(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ffu;
  if (exp == 0) {
    const float v = ldexpf((float)man, -24);
    return sign ? -v : v;
  }
  const uint32_t b = sign | ((exp == 31 ? 255u : exp + 112u) << 23) | (man << 13);
  float f;
  memcpy(&f, &b, 4);
  return f;
}

static int engine_build(rb200_engine* e, const rb200_engine_config* cfg);
static int alloc_workspace(rb200_engine* e, int max_batch, int max_beams, int max_src_len);
static void free_workspace(rb200_engine* e);

extern "C" {

int rb200_engine_create(const rb200_engine_config* cfg, rb200_engine** out) {
  RB_REQUIRE(cfg && out, "null argument");
  RB_REQUIRE(cfg->d_kv == 64, "d_kv must be 64 (t5-base/large), got %d", cfg->d_kv);
  RB_REQUIRE(cfg->precision >= 0 && cfg->precision <= 5, "unknown precision %d", cfg->precision);
  RB_REQUIRE(cfg->d_model % 8 == 0 && cfg->d_ff % 8 == 0, "d_model and d_ff must be multiples of 8");
  RB_REQUIRE(cfg->decoder_vocab_size % 4 == 0, "decoder_vocab_size must be a multiple of 4");
  RB_REQUIRE(cfg->max_batch >= 1 && cfg->max_beams >= 1 && cfg->max_src_len >= 1 && cfg->docid_len >= 1,
             "max_batch, max_beams, max_src_len, docid_len must be >= 1");
  int ndev = 0;
  RB_CUDA(cudaGetDeviceCount(&ndev));
  RB_REQUIRE(cfg->device >= 0 && cfg->device < ndev, "device %d not present (%d devices)", cfg->device, ndev);
  RB_CUDA(cudaSetDevice(cfg->device));
  rb200_engine* e = new (std::nothrow) rb200_engine();
  if (!e) return rb::fail(RB200_ERR_NOMEM, "out of memory");
  const int st = engine_build(e, cfg);
  if (st != 0) {                      // e.g. out of device memory half way: give everything back (keeps last_error)
    rb200_engine_free(e);
    return st;
  }
  *out = e;
  return 0;
}

}  // extern "C"

// allocations of rb200_engine_create; on failure the caller frees whatever was allocated so far
static int engine_build(rb200_engine* e, const rb200_engine_config* cfg) {
  e->cfg = *cfg;
  e->mode = cfg->precision;
  e->planes = rb::prec_planes(e->mode);
  e->elem = rb::prec_elem_bytes(e->mode);
  e->d = cfg->d_model; e->H = cfg->num_heads; e->inner = cfg->num_heads * cfg->d_kv; e->dff = cfg->d_ff;
  e->V = cfg->decoder_vocab_size; e->Lmodel = cfg->docid_len;
  const int d = e->d, inner = e->inner, dff = e->dff;
  e->enc.resize(cfg->num_layers);
  e->dec.resize(cfg->num_decoder_layers);
  for (auto& l : e->enc) {
    RB_TRY(alloc_packed(e, &l.qkv, 3 * inner, d));
    RB_TRY(alloc_packed(e, &l.o, d, inner));
    RB_TRY(alloc_packed(e, &l.wi, dff, d));
    RB_TRY(alloc_packed(e, &l.wo, d, dff));
    RB_TRY(dev_alloc(e, (void**)&l.ln0, d * 4));
    RB_TRY(dev_alloc(e, (void**)&l.ln1, d * 4));
  }
  for (auto& l : e->dec) {
    RB_TRY(alloc_packed(e, &l.qkv, 3 * inner, d));
    RB_TRY(alloc_packed(e, &l.o, d, inner));
    RB_TRY(alloc_packed(e, &l.cq, inner, d));
    RB_TRY(alloc_packed(e, &l.co, d, inner));
    RB_TRY(alloc_packed(e, &l.wi, dff, d));
    RB_TRY(alloc_packed(e, &l.wo, d, dff));
    RB_TRY(dev_alloc(e, (void**)&l.ln0, d * 4));
    RB_TRY(dev_alloc(e, (void**)&l.ln1, d * 4));
    RB_TRY(dev_alloc(e, (void**)&l.ln2, d * 4));
  }
  RB_TRY(alloc_packed(e, &e->ckv, (int64_t)cfg->num_decoder_layers * 2 * inner, d));
  e->out_tab.resize(e->Lmodel);
  e->in_tab.resize(e->Lmodel, nullptr);
  for (int t = 0; t < e->Lmodel; ++t) {
    RB_TRY(alloc_packed(e, &e->out_tab[t], e->V, d));
    RB_TRY(dev_alloc(e, (void**)&e->in_tab[t], (int64_t)e->V * d * 4));
  }
  RB_TRY(dev_alloc(e, (void**)&e->shared_emb, (int64_t)cfg->vocab_size * d * 4));
  RB_TRY(dev_alloc(e, (void**)&e->start_emb, d * 4));
  RB_TRY(dev_alloc(e, (void**)&e->enc_final_ln, d * 4));
  RB_TRY(dev_alloc(e, (void**)&e->dec_final_ln, d * 4));
  {
    const char* f = getenv("RB200_FOLD");
    // NormFold (default; RB200_FOLD=0 turns it off): the T5 layer norms live in the epilogues of the GEMMs around
    // them instead of 3 RMSNorm launches per layer. It pays since the new residual values leave the epilogue as TMA
    // stores (58.1 -> 56.6 ms per batch at the bench shape); with row-per-lane 16-byte stores it cost as much as the
    // RMSNorm launches it replaced.
    e->fold = e->mode != RB200_PREC_FP32 && d % 64 == 0 && !(f && f[0] == '0');
    e->np = d / 64;
  }
  {
    const char* tl = getenv("RB200_TAIL");
    e->tail = !(tl && tl[0] == '0');
    RB_CUDA(cudaMallocHost((void**)&e->flag_host, 2 * sizeof(int32_t)));
  }
  RB_TRY(dev_alloc(e, (void**)&e->overflow, 2 * 4));
  RB_CUDA(cudaMemset(e->overflow, 0, 8));
  return alloc_workspace(e, cfg->max_batch, cfg->max_beams, cfg->max_src_len);
}

// Everything whose size depends on (max_batch, max_beams, max_src_len): activations, KV cache, beam state, host-call
// staging. Kept apart from the packed weights so that rb200_engine_resize never touches those.
static void free_workspace(rb200_engine* e) {
  Lane& l = e->lanes[0];
  void* lb[] = {l.x, l.x_full, l.xn, l.qkv, l.q2, l.ctx, l.hbuf, l.logits, l.cross_kv, l.enc_out, l.cache_k, l.cache_v,
                l.ss[0], l.ss[1], e->ids_dev, e->mask_dev, e->seq_dev, e->score_dev, e->leaf_dev, e->enc_bias};
  for (void* b : lb) cudaFree(b);
  rb200_beam_free(l.beam);
  l = Lane();
  e->ids_dev = e->mask_dev = e->seq_dev = nullptr;
  e->score_dev = nullptr; e->leaf_dev = nullptr; e->enc_bias = nullptr; e->enc_bias_S = 0;
  e->ws_bytes -= e->ws_only_bytes;
  e->ws_only_bytes = 0;
}

static int alloc_workspace(rb200_engine* e, int max_batch, int max_beams, int max_src_len) {
  const int d = e->d, inner = e->inner, dff = e->dff;
  const auto& cfg = e->cfg;
  const int64_t before = e->ws_bytes;
  const int64_t pe = (int64_t)e->planes * e->elem;
  Lane& l = e->lanes[0];
  l.Rcap = (int64_t)max_batch * max_beams;
  l.BScap = (int64_t)max_batch * max_src_len;
  // the forced tail runs up to min(Lmodel, RB_TAIL_MAX_L) positions of every beam row in one pass
  const int tail_pos = e->Lmodel < RB_TAIL_MAX_L ? e->Lmodel : RB_TAIL_MAX_L;
  l.Mcap = std::max(l.Rcap * tail_pos, l.BScap);
  RB_TRY(dev_alloc(e, (void**)&l.x, l.Mcap * d * 4));
  RB_TRY(dev_alloc(e, (void**)&l.x_full, l.Rcap * d * 4));
  RB_TRY(dev_alloc(e, &l.xn, l.Mcap * d * pe));
  RB_TRY(dev_alloc(e, (void**)&l.qkv, l.Mcap * 3 * inner * 4));
  RB_TRY(dev_alloc(e, (void**)&l.q2, l.Mcap * inner * 4));
  RB_TRY(dev_alloc(e, &l.ctx, l.Mcap * inner * pe));
  RB_TRY(dev_alloc(e, &l.hbuf, l.Mcap * dff * pe));
  RB_TRY(dev_alloc(e, (void**)&l.logits, l.Mcap * e->V * 4));
  RB_TRY(dev_alloc(e, (void**)&l.cross_kv, l.BScap * cfg.num_decoder_layers * 2 * inner * 4));
  RB_TRY(dev_alloc(e, (void**)&l.enc_out, l.BScap * d * 4));
  const int64_t cache = (int64_t)cfg.num_decoder_layers * e->Lmodel * l.Rcap * inner * 4;
  RB_TRY(dev_alloc(e, (void**)&l.cache_k, cache));
  RB_TRY(dev_alloc(e, (void**)&l.cache_v, cache));
  for (int b = 0; b < 2; ++b) RB_TRY(dev_alloc(e, (void**)&l.ss[b], l.Mcap * std::max(e->np, 1) * 4));
  // planes of the activation buffers may be read past M by TMA boxes: start from defined contents
  RB_CUDA(cudaMemset(l.xn, 0, (size_t)(l.Mcap * d * pe)));
  RB_CUDA(cudaMemset(l.ctx, 0, (size_t)(l.Mcap * inner * pe)));
  RB_CUDA(cudaMemset(l.hbuf, 0, (size_t)(l.Mcap * dff * pe)));
  RB_TRY(rb200_beam_create(cfg.device, max_batch, max_beams, e->Lmodel, e->V, &l.beam));
  RB_TRY(dev_alloc(e, (void**)&e->ids_dev, l.BScap * 8));
  RB_TRY(dev_alloc(e, (void**)&e->mask_dev, l.BScap * 8));
  RB_TRY(dev_alloc(e, (void**)&e->seq_dev, l.Rcap * (e->Lmodel + 1) * 8));
  RB_TRY(dev_alloc(e, (void**)&e->score_dev, l.Rcap * 4));
  RB_TRY(dev_alloc(e, (void**)&e->leaf_dev, l.Rcap * 2 * 4));
  e->ws_batch = max_batch; e->ws_beams = max_beams; e->ws_src = max_src_len;
  e->cfg.max_batch = max_batch; e->cfg.max_beams = max_beams; e->cfg.max_src_len = max_src_len;
  e->ws_only_bytes = e->ws_bytes - before;
  return 0;
}

extern "C" {

int rb200_engine_free(rb200_engine* e) {
  if (!e) return 0;
  auto fp = [](Packed& p) { cudaFree(p.ptr); };
  for (auto& l : e->enc) { fp(l.qkv); fp(l.o); fp(l.wi); fp(l.wo); cudaFree(l.ln0); cudaFree(l.ln1); }
  for (auto& l : e->dec) {
    fp(l.qkv); fp(l.o); fp(l.cq); fp(l.co); fp(l.wi); fp(l.wo);
    cudaFree(l.ln0); cudaFree(l.ln1); cudaFree(l.ln2);
  }
  fp(e->ckv);
  for (auto& p : e->out_tab) fp(p);
  for (auto p : e->in_tab) cudaFree(p);
  void* bufs[] = {e->shared_emb, e->start_emb, e->enc_final_ln, e->dec_final_ln, e->dec_bias, e->overflow};
  for (void* b : bufs) cudaFree(b);
  free_workspace(e);
  cudaFree(e->in_tab_dev);
  if (e->flag_host) cudaFreeHost(e->flag_host);
  for (auto& st : e->stash) cudaFree(st.src);
  for (auto ev : e->events) cudaEventDestroy(ev);
  delete e;
  return 0;
}

int64_t rb200_engine_workspace_bytes(const rb200_engine* e) { return e ? e->ws_bytes : 0; }

int rb200_engine_set_weight(rb200_engine* e, const char* name_c, const float* data, int64_t numel, void* stream) {
  RB_REQUIRE(e && name_c && data, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  const std::string name(name_c);
  const int64_t d = e->d, inner = e->inner, dff = e->dff, V = e->V;
  auto expect = [&](int64_t n) -> int {
    if (numel != n) return rb::fail(RB200_ERR_INVALID, "%s: expected %lld elements, got %lld", name_c, (long long)n,
                                    (long long)numel);
    return 0;
  };
  int layer = -1, sub = -1, pos = -1;
  char what[64] = {0};
  int st = 0;
  const float out_scale_fold = e->cfg.scaleup_output_hidden ? 1.0f / sqrtf((float)d) : 1.0f;
  if (name == "shared.weight" || name == "encoder.embed_tokens.weight") {
    RB_TRY(expect((int64_t)e->cfg.vocab_size * d));
    st = copy_f32(e->shared_emb, data, numel, s);
    e->have.insert("shared.weight");
    return st;
  } else if (name == "start_token_embed") {
    RB_TRY(expect(d));
    st = copy_f32(e->start_emb, data, numel, s);
  } else if (name == "encoder.final_layer_norm.weight") {
    RB_TRY(expect(d));
    st = copy_f32(e->enc_final_ln, data, numel, s);
  } else if (name == "decoder.final_layer_norm.weight") {
    RB_TRY(expect(d));
    st = copy_f32(e->dec_final_ln, data, numel, s);
  } else if (sscanf(name_c, "list_decoder_embeds.%d.weight", &pos) == 1) {
    RB_REQUIRE(pos >= 0 && pos < e->Lmodel, "%s: position outside [0, %d)", name_c, e->Lmodel);
    RB_TRY(expect(V * d));
    RB_TRY(copy_f32(e->in_tab[pos], data, numel, s));
    if (e->cfg.shared_output_input_embeds)
      st = pack_rows(e, &e->out_tab[pos], 0, data, V, s, e->dec_final_ln, out_scale_fold);
  } else if (sscanf(name_c, "list_output_embeds.%d.weight", &pos) == 1) {
    RB_REQUIRE(pos >= 0 && pos < e->Lmodel, "%s: position outside [0, %d)", name_c, e->Lmodel);
    RB_TRY(expect(V * d));
    if (!e->cfg.shared_output_input_embeds)
      st = pack_rows(e, &e->out_tab[pos], 0, data, V, s, e->dec_final_ln, out_scale_fold);
  } else if (sscanf(name_c, "encoder.block.%d.layer.%d.%63s", &layer, &sub, what) == 3 ||
             sscanf(name_c, "decoder.block.%d.layer.%d.%63s", &layer, &sub, what) == 3) {
    const bool is_dec = name.compare(0, 7, "decoder") == 0;
    auto& layers = is_dec ? e->dec : e->enc;
    RB_REQUIRE(layer >= 0 && layer < (int)layers.size(), "%s: no such block", name_c);
    Layer& l = layers[layer];
    const std::string w(what);
    const int ff_sub = is_dec ? 2 : 1;
    if (w == "layer_norm.weight") {
      RB_TRY(expect(d));
      float* dst = sub == 0 ? l.ln0 : (sub == 1 ? l.ln1 : l.ln2);
      RB_REQUIRE(dst != nullptr && sub <= ff_sub, "%s: no such sub-layer", name_c);
      st = copy_f32(dst, data, numel, s);
    } else if (sub == 0 && w == "SelfAttention.relative_attention_bias.weight") {
      RB_TRY(expect((int64_t)e->cfg.num_buckets * e->H));
      RB_REQUIRE(layer == 0, "%s: only block 0 carries the relative attention bias", name_c);
      auto& host = is_dec ? e->dec_rel_host : e->enc_rel_host;
      host.resize(numel);
      RB_CUDA(cudaMemcpyAsync(host.data(), data, numel * 4, cudaMemcpyDeviceToHost, s));
      RB_CUDA(cudaStreamSynchronize(s));
    } else if (sub == 0 && (w == "SelfAttention.q.weight" || w == "SelfAttention.k.weight" ||
                            w == "SelfAttention.v.weight")) {
      RB_TRY(expect(inner * d));
      const int which = w[14] == 'q' ? 0 : (w[14] == 'k' ? 1 : 2);
      st = pack_rows(e, &l.qkv, which * inner, data, inner, s, l.ln0);
    } else if (sub == 0 && w == "SelfAttention.o.weight") {
      RB_TRY(expect(d * inner));
      st = pack_rows(e, &l.o, 0, data, d, s);
    } else if (is_dec && sub == 1 && w == "EncDecAttention.q.weight") {
      RB_TRY(expect(inner * d));
      st = pack_rows(e, &l.cq, 0, data, inner, s, l.ln1);
    } else if (is_dec && sub == 1 && (w == "EncDecAttention.k.weight" || w == "EncDecAttention.v.weight")) {
      RB_TRY(expect(inner * d));
      const int which = w[16] == 'k' ? 0 : 1;
      st = pack_rows(e, &e->ckv, ((int64_t)layer * 2 + which) * inner, data, inner, s, e->enc_final_ln);
    } else if (is_dec && sub == 1 && w == "EncDecAttention.o.weight") {
      RB_TRY(expect(d * inner));
      st = pack_rows(e, &l.co, 0, data, d, s);
    } else if (sub == ff_sub && w == "DenseReluDense.wi.weight") {
      RB_TRY(expect(dff * d));
      st = pack_rows(e, &l.wi, 0, data, dff, s, is_dec ? l.ln2 : l.ln1);
    } else if (sub == ff_sub && w == "DenseReluDense.wo.weight") {
      RB_TRY(expect(d * dff));
      st = pack_rows(e, &l.wo, 0, data, d, s);
    } else {
      return rb::fail(RB200_ERR_INVALID, "unknown weight %s", name_c);
    }
  } else {
    return rb::fail(RB200_ERR_INVALID, "unknown weight %s", name_c);
  }
  if (st == 0) e->have.insert(name);
  return st;
}

int rb200_engine_finalize_weights(rb200_engine* e, void* stream) {
  RB_REQUIRE(e, "null argument");
  std::vector<std::string> need = {"shared.weight", "start_token_embed", "encoder.final_layer_norm.weight",
                                   "decoder.final_layer_norm.weight",
                                   "encoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight",
                                   "decoder.block.0.layer.0.SelfAttention.relative_attention_bias.weight"};
  for (int side = 0; side < 2; ++side) {
    const int nl = side ? e->cfg.num_decoder_layers : e->cfg.num_layers;
    const std::string pre = side ? "decoder.block." : "encoder.block.";
    for (int i = 0; i < nl; ++i) {
      const std::string p = pre + std::to_string(i) + ".layer.";
      for (const char* w : {"q", "k", "v", "o"}) need.push_back(p + "0.SelfAttention." + w + ".weight");
      need.push_back(p + "0.layer_norm.weight");
      int ff = 1;
      if (side) {
        for (const char* w : {"q", "k", "v", "o"}) need.push_back(p + "1.EncDecAttention." + w + ".weight");
        need.push_back(p + "1.layer_norm.weight");
        ff = 2;
      }
      need.push_back(p + std::to_string(ff) + ".DenseReluDense.wi.weight");
      need.push_back(p + std::to_string(ff) + ".DenseReluDense.wo.weight");
      need.push_back(p + std::to_string(ff) + ".layer_norm.weight");
    }
  }
  for (int t = 0; t < e->Lmodel; ++t) {
    need.push_back("list_decoder_embeds." + std::to_string(t) + ".weight");
    if (!e->cfg.shared_output_input_embeds) need.push_back("list_output_embeds." + std::to_string(t) + ".weight");
  }
  for (const auto& n : need)
    if (!e->have.count(n)) return rb::fail(RB200_ERR_STATE, "weight %s has not been set", n.c_str());
  if (e->in_tab_dev == nullptr) {
    RB_CUDA(cudaMalloc((void**)&e->in_tab_dev, e->in_tab.size() * sizeof(float*)));
    RB_CUDA(cudaMemcpy(e->in_tab_dev, e->in_tab.data(), e->in_tab.size() * sizeof(float*), cudaMemcpyHostToDevice));
  }
  // NormFold: every layer-norm vector is on the device now - pack W * diag(ln) for the matrices that follow one
  for (auto& st : e->stash) {
    char* dst = static_cast<char*>(st.w->ptr) + st.row0 * st.w->K * e->elem;
    const float scale = (rb::prec_is_fp16(e->mode) ? rb::kFp16WeightScale : 1.0f) * st.scale;
    RB_TRY(rb::launch_pack_planes_cols(st.src, dst, st.rows * st.w->K, st.w->plane, e->mode, scale, st.ln, st.w->K,
                                       e->overflow + 1, (cudaStream_t)stream));
  }
  RB_TRY(build_bias_tables(e, 0, (cudaStream_t)stream));
  RB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  for (auto& st : e->stash) cudaFree(st.src);
  e->stash.clear();
  int wflag = 0;
  RB_CUDA(cudaMemcpy(&wflag, e->overflow + 1, 4, cudaMemcpyDeviceToHost));
  if (wflag)
    return rb::fail(RB200_ERR_INVALID, "a weight exceeds the fp16 range after the 2^8 pre-scale: use tf32x3");
  e->finalized = true;
  return 0;
}

}  // extern "C"

namespace {

int lane_encode(rb200_engine* e, Lane& l, const int64_t* ids, const int64_t* mask, int batch, int S, int num_beams,
                cudaStream_t s) {
  l.B = batch; l.S = S; l.nb = num_beams; l.cur_mask = mask;
  const int64_t rows = (int64_t)batch * S;
  const int d = e->d, inner = e->inner, dff = e->dff;
  const float eps = e->cfg.layer_norm_eps;
  RB_TRY(rb::launch_embed_rows(e->shared_emb, ids, l.x, rows, d, s));
  // bidirectional self-attention = the decode cross-attention kernel with the S rows of a sequence as its "beams",
  // K/V taken from the fused q|k|v rows, plus the relative position bias
  auto enc_attn = [&]() {
    rb::CrossAttnArgs ca;
    ca.q = l.qkv; ca.q_ld = 3 * inner; ca.kv = l.qkv; ca.ld = 3 * inner; ca.k_off = inner; ca.v_off = 2 * inner;
    ca.mask = mask; ca.M = (int)rows; ca.H = e->H; ca.S = S; ca.rows_per_query = S; ca.rel_bias = e->enc_bias;
    ca.rel_S = e->enc_bias_S;
    return rb::launch_cross_attn_decode(ca, e->act(l, l.ctx, inner), s);
  };
  if (e->fold) {
    // NormFold chain: p = parity of the partial-sum table that holds the CURRENT norm point
    int p = 0;
    auto consumer = [&]() {
      rb::NormFold nf;
      nf.ss_prev = l.ss[p ^ 1]; nf.ss_cur = l.ss[p]; nf.np = e->np; nf.inv_d = 1.0f / (float)d; nf.eps = eps;
      nf.scaled = true;
      return nf;
    };
    auto producer = [&]() {
      rb::NormFold nf;
      nf.ss_prev = l.ss[p]; nf.ss_out = l.ss[p ^ 1]; nf.np = e->np; nf.inv_d = 1.0f / (float)d; nf.eps = eps;
      p ^= 1;
      return nf;
    };
    RB_TRY(rb::launch_norm_init(l.x, e->act(l, l.xn, d), l.ss[0], l.ss[1], e->np, rows, d, eps, s));
    for (auto& w : e->enc) {
      RB_TRY(gemm(e, l, l.xn, d, w.qkv, l.qkv, 3 * inner, ActOut{}, rows, rb::EPI_STORE, s, consumer()));
      RB_TRY(enc_attn());
      RB_TRY(gemm(e, l, l.ctx, inner, w.o, l.x, d, e->act(l, l.xn, d), rows, rb::EPI_RESID_NORM, s, producer()));
      RB_TRY(gemm(e, l, l.xn, d, w.wi, nullptr, 0, e->act(l, l.hbuf, dff), rows, rb::EPI_RELU_ACT, s, consumer()));
      RB_TRY(gemm(e, l, l.hbuf, dff, w.wo, l.x, d, e->act(l, l.xn, d), rows, rb::EPI_RESID_NORM, s, producer()));
    }
    RB_TRY(rb::launch_rmsnorm_f32(l.x, e->enc_final_ln, l.enc_out, rows, d, eps, s));
    for (size_t i = 0; i < e->dec.size(); ++i)
      RB_TRY(gemm(e, l, l.xn, d, sub_rows(e, e->ckv, (int64_t)i * 2 * inner, 2 * inner),
                  l.cross_kv + (int64_t)i * l.BScap * 2 * inner, 2 * inner, ActOut{}, rows, rb::EPI_STORE, s, consumer()));
    return 0;
  }
  for (auto& w : e->enc) {
    RB_TRY(rb::launch_rmsnorm(l.x, w.ln0, e->act(l, l.xn, d), rows, d, eps, 1.0f, s));
    RB_TRY(gemm(e, l, l.xn, d, w.qkv, l.qkv, 3 * inner, ActOut{}, rows, rb::EPI_STORE, s));
    RB_TRY(enc_attn());
    RB_TRY(gemm(e, l, l.ctx, inner, w.o, l.x, d, ActOut{}, rows, rb::EPI_RESIDUAL, s));
    RB_TRY(rb::launch_rmsnorm(l.x, w.ln1, e->act(l, l.xn, d), rows, d, eps, 1.0f, s));
    RB_TRY(gemm(e, l, l.xn, d, w.wi, nullptr, 0, e->act(l, l.hbuf, dff), rows, rb::EPI_RELU_ACT, s));
    RB_TRY(gemm(e, l, l.hbuf, dff, w.wo, l.x, d, ActOut{}, rows, rb::EPI_RESIDUAL, s));
  }
  RB_TRY(rb::launch_rmsnorm_f32(l.x, e->enc_final_ln, l.enc_out, rows, d, eps, s));
  RB_TRY(rb::launch_rmsnorm(l.x, e->enc_final_ln, e->act(l, l.xn, d), rows, d, eps, 1.0f, s));
  // cross-attention K/V of every decoder layer, one GEMM per layer into a layer-major table [Nl][B*S][K | V]: the
  // rows a decode-step cross-attention launch reads are then one dense 6 KB-per-position block per query instead of
  // 3 KB pieces 74 KB apart (DRAM page locality)
  for (size_t i = 0; i < e->dec.size(); ++i)
    RB_TRY(gemm(e, l, l.xn, d, sub_rows(e, e->ckv, (int64_t)i * 2 * inner, 2 * inner),
                l.cross_kv + (int64_t)i * l.BScap * 2 * inner, 2 * inner, ActOut{}, rows, rb::EPI_STORE, s));
  return 0;
}

int lane_decode_step(rb200_engine* e, Lane& l, const rb200_beam* beam, int t, float* logits, cudaStream_t s) {
  const int rpq = (t == 0) ? 1 : l.nb;
  const int64_t M = (int64_t)beam->n_active * rpq;          // rows of the queries still stepping, compact order
  const int32_t* qlist = beam->compacted ? beam->qlist : nullptr;
  const int d = e->d, inner = e->inner, dff = e->dff;
  const float eps = e->cfg.layer_norm_eps;
  if (M == 0) return 0;
  if (t == 0) {
    RB_TRY(rb::launch_broadcast_row(e->start_emb, l.x, M, d, s));
  } else {
    // the previous beam step left the inputs at original row ids: pull the stepping queries' rows together
    RB_TRY(rb::launch_gather_rows(beam, l.x_full, l.x, d, s));
  }
  const int64_t layer_cache = (int64_t)e->Lmodel * l.Rcap * inner;
  auto self_attn = [&](size_t i) {
    rb::SelfAttnArgs sa;
    sa.qkv = l.qkv; sa.cache_k = l.cache_k + i * layer_cache; sa.cache_v = l.cache_v + i * layer_cache;
    sa.anc = beam->anc[beam->cur]; sa.bias = e->dec_bias; sa.row_cap = l.Rcap;
    sa.M = (int)M; sa.H = e->H; sa.L = e->Lmodel; sa.t = t; sa.rpq = rpq; sa.nb = l.nb; sa.qlist = qlist;
    return rb::launch_self_attn_decode(sa, e->act(l, l.ctx, inner), s);
  };
  auto cross_attn = [&](size_t i) {
    rb::CrossAttnArgs ca;
    ca.q = l.q2; ca.kv = l.cross_kv + (int64_t)i * l.BScap * 2 * inner; ca.ld = 2 * inner; ca.k_off = 0;
    ca.v_off = inner; ca.mask = l.cur_mask; ca.M = (int)M; ca.H = e->H; ca.S = l.S;
    ca.rows_per_query = rpq; ca.qmap = qlist;
    return rb::launch_cross_attn_decode(ca, e->act(l, l.ctx, inner), s);
  };
  if (e->fold) {
    int p = 0;   // parity of the partial-sum table holding the current norm point (see lane_encode)
    auto consumer = [&]() {
      rb::NormFold nf;
      nf.ss_prev = l.ss[p ^ 1]; nf.ss_cur = l.ss[p]; nf.np = e->np; nf.inv_d = 1.0f / (float)d; nf.eps = eps;
      nf.scaled = true;
      return nf;
    };
    auto producer = [&]() {
      rb::NormFold nf;
      nf.ss_prev = l.ss[p]; nf.ss_out = l.ss[p ^ 1]; nf.np = e->np; nf.inv_d = 1.0f / (float)d; nf.eps = eps;
      p ^= 1;
      return nf;
    };
    RB_TRY(rb::launch_norm_init(l.x, e->act(l, l.xn, d), l.ss[0], l.ss[1], e->np, M, d, eps, s));
    for (size_t i = 0; i < e->dec.size(); ++i) {
      Layer& w = e->dec[i];
      RB_TRY(gemm(e, l, l.xn, d, w.qkv, l.qkv, 3 * inner, ActOut{}, M, rb::EPI_STORE, s, consumer()));
      RB_TRY(self_attn(i));
      RB_TRY(gemm(e, l, l.ctx, inner, w.o, l.x, d, e->act(l, l.xn, d), M, rb::EPI_RESID_NORM, s, producer()));
      RB_TRY(gemm(e, l, l.xn, d, w.cq, l.q2, inner, ActOut{}, M, rb::EPI_STORE, s, consumer()));
      RB_TRY(cross_attn(i));
      RB_TRY(gemm(e, l, l.ctx, inner, w.co, l.x, d, e->act(l, l.xn, d), M, rb::EPI_RESID_NORM, s, producer()));
      RB_TRY(gemm(e, l, l.xn, d, w.wi, nullptr, 0, e->act(l, l.hbuf, dff), M, rb::EPI_RELU_ACT, s, consumer()));
      RB_TRY(gemm(e, l, l.hbuf, dff, w.wo, l.x, d, e->act(l, l.xn, d), M, rb::EPI_RESID_NORM, s, producer()));
    }
    // final layer norm (and the optional d^-1/2) live in the packed output tables
    RB_TRY(gemm(e, l, l.xn, d, e->out_tab[t], logits, e->V, ActOut{}, M, rb::EPI_STORE, s, consumer()));
    return 0;
  }
  for (size_t i = 0; i < e->dec.size(); ++i) {
    Layer& w = e->dec[i];
    RB_TRY(rb::launch_rmsnorm(l.x, w.ln0, e->act(l, l.xn, d), M, d, eps, 1.0f, s));
    RB_TRY(gemm(e, l, l.xn, d, w.qkv, l.qkv, 3 * inner, ActOut{}, M, rb::EPI_STORE, s));
    RB_TRY(self_attn(i));
    RB_TRY(gemm(e, l, l.ctx, inner, w.o, l.x, d, ActOut{}, M, rb::EPI_RESIDUAL, s));
    RB_TRY(rb::launch_rmsnorm(l.x, w.ln1, e->act(l, l.xn, d), M, d, eps, 1.0f, s));
    RB_TRY(gemm(e, l, l.xn, d, w.cq, l.q2, inner, ActOut{}, M, rb::EPI_STORE, s));
    RB_TRY(cross_attn(i));
    RB_TRY(gemm(e, l, l.ctx, inner, w.co, l.x, d, ActOut{}, M, rb::EPI_RESIDUAL, s));
    RB_TRY(rb::launch_rmsnorm(l.x, w.ln2, e->act(l, l.xn, d), M, d, eps, 1.0f, s));
    RB_TRY(gemm(e, l, l.xn, d, w.wi, nullptr, 0, e->act(l, l.hbuf, dff), M, rb::EPI_RELU_ACT, s));
    RB_TRY(gemm(e, l, l.hbuf, dff, w.wo, l.x, d, ActOut{}, M, rb::EPI_RESIDUAL, s));
  }
  const float scale = e->cfg.scaleup_output_hidden ? 1.0f / sqrtf((float)d) : 1.0f;
  RB_TRY(rb::launch_rmsnorm(l.x, e->dec_final_ln, e->act(l, l.xn, d), M, d, eps, scale, s));
  RB_TRY(gemm(e, l, l.xn, d, e->out_tab[t], logits, e->V, ActOut{}, M, rb::EPI_STORE, s));
  return 0;
}

// Forced tail: the remaining positions of every frozen query's beams in ONE teacher-forced pass (rows position-major
// and ragged, see rb::TailLayout). Same kernels and per-row arithmetic as the decoder steps they replace; what changes
// is that the GEMMs see all (row, position) pairs at once, a lineage's K/V are read once for all of its positions, and
// ~130 launches replace ~130 per step. With trie == nullptr the tokens come from the beam state's forced history
// (rb200_engine_forward: teacher forcing from position 0, no cached prefix).
int lane_tail(rb200_engine* e, Lane& l, const rb200_trie* trie, const rb::TailLayout& lay, float* hidden_out,
              cudaStream_t s) {
  rb200_beam* beam = l.beam;
  const int nfz_rows = beam->n_frozen * beam->nb;
  const int64_t M = lay.off[lay.P];
  const int d = e->d, inner = e->inner, dff = e->dff;
  const float eps = e->cfg.layer_norm_eps;
  if (M == 0) return 0;
  RB_REQUIRE(M <= l.Mcap, "forced tail of %lld rows exceeds the workspace capacity %lld", (long long)M, (long long)l.Mcap);
  RB_TRY(rb::launch_tail_prepare(beam, trie, lay, e->in_tab_dev, e->start_emb, l.x, d, s));
  const int32_t* qstart = trie ? beam->qstate : nullptr;
  const bool self_planes = rb::tail_self_attn_reads_planes(e->mode);
  const int64_t layer_cache = (int64_t)e->Lmodel * l.Rcap * inner;
  // NormFold chain (RB200_FOLD=1): the layer norms live in the GEMM epilogues around them, see lane_encode
  int fp = 0;
  auto consumer = [&](int64_t row0 = 0) {
    rb::NormFold nf;
    if (!e->fold) return nf;
    nf.ss_prev = l.ss[fp ^ 1] + row0 * e->np; nf.ss_cur = l.ss[fp] + row0 * e->np; nf.np = e->np;
    nf.inv_d = 1.0f / (float)d; nf.eps = eps; nf.scaled = true;
    return nf;
  };
  auto producer = [&]() {
    rb::NormFold nf;
    nf.ss_prev = l.ss[fp]; nf.ss_out = l.ss[fp ^ 1]; nf.np = e->np; nf.inv_d = 1.0f / (float)d; nf.eps = eps;
    fp ^= 1;
    return nf;
  };
  // residual add (+ the following layer norm when folded, else a separate RMSNorm launch before the next GEMM)
  auto residual = [&](const void* A, int64_t a_len, const Packed& w) {
    if (e->fold) return gemm(e, l, A, a_len, w, l.x, d, e->act(l, l.xn, d), M, rb::EPI_RESID_NORM, s, producer());
    return gemm(e, l, A, a_len, w, l.x, d, ActOut{}, M, rb::EPI_RESIDUAL, s);
  };
  auto norm = [&](const float* ln) {
    if (e->fold) return 0;
    return rb::launch_rmsnorm(l.x, ln, e->act(l, l.xn, d), M, d, eps, 1.0f, s);
  };
  if (e->fold) RB_TRY(rb::launch_norm_init(l.x, e->act(l, l.xn, d), l.ss[0], l.ss[1], e->np, M, d, eps, s));
  for (size_t i = 0; i < e->dec.size(); ++i) {
    Layer& w = e->dec[i];
    RB_TRY(norm(w.ln0));
    rb::TailAttnArgs ta;
    if (self_planes) {     // q | k | v leave the GEMM as fp16 hi/lo planes (same bytes as the fp32 buffer they replace)
      RB_TRY(gemm(e, l, l.xn, d, w.qkv, nullptr, 0, e->act(l, l.qkv, 3 * inner), M, rb::EPI_PLANES, s, consumer()));
      ta.qkv_hi = reinterpret_cast<const __half*>(l.qkv);
      ta.qkv_plane = l.Mcap * 3 * inner;
    } else {
      RB_TRY(gemm(e, l, l.xn, d, w.qkv, l.qkv, 3 * inner, ActOut{}, M, rb::EPI_STORE, s, consumer()));
    }
    ta.qkv = l.qkv; ta.cache_k = l.cache_k + i * layer_cache; ta.cache_v = l.cache_v + i * layer_cache;
    ta.anc = beam->fz_anc; ta.bias = e->dec_bias; ta.row_cap = l.Rcap;
    ta.R = nfz_rows; ta.H = e->H; ta.L = e->Lmodel; ta.nb = beam->nb;
    ta.fz_list = beam->fz_list; ta.qstart = qstart; ta.lay = lay;
    RB_TRY(rb::launch_self_attn_tail(ta, e->act(l, l.ctx, inner), s));
    RB_TRY(residual(l.ctx, inner, w.o));
    RB_TRY(norm(w.ln1));
    RB_TRY(gemm(e, l, l.xn, d, w.cq, l.q2, inner, ActOut{}, M, rb::EPI_STORE, s, consumer()));
    rb::CrossAttnArgs ca;
    ca.q = l.q2; ca.kv = l.cross_kv + (int64_t)i * l.BScap * 2 * inner; ca.ld = 2 * inner; ca.k_off = 0;
    ca.v_off = inner; ca.mask = l.cur_mask; ca.M = nfz_rows; ca.H = e->H; ca.S = l.S; ca.rows_per_query = beam->nb;
    ca.qmap = beam->fz_list; ca.ragged = 1; ca.qstart = qstart; ca.lay = lay;
    RB_TRY(rb::launch_cross_attn_decode(ca, e->act(l, l.ctx, inner), s));
    RB_TRY(residual(l.ctx, inner, w.co));
    RB_TRY(norm(w.ln2));
    RB_TRY(gemm(e, l, l.xn, d, w.wi, nullptr, 0, e->act(l, l.hbuf, dff), M, rb::EPI_RELU_ACT, s, consumer()));
    RB_TRY(residual(l.hbuf, dff, w.wo));
  }
  const float scale = e->cfg.scaleup_output_hidden ? 1.0f / sqrtf((float)d) : 1.0f;
  // (folded: the final layer norm and the optional d^-1/2 live in the packed output tables)
  if (!e->fold) RB_TRY(rb::launch_rmsnorm(l.x, e->dec_final_ln, e->act(l, l.xn, d), M, d, eps, scale, s));
  if (hidden_out) RB_TRY(rb::launch_rmsnorm_f32(l.x, e->dec_final_ln, hidden_out, M, d, eps, s, scale));
  // LM head: the output table differs per position -> one GEMM per position block
  for (int p = 0; p < lay.P; ++p) {
    const int64_t n_p = lay.off[p + 1] - lay.off[p];
    if (n_p == 0) continue;
    RB_TRY(gemm(e, l, static_cast<const char*>(l.xn) + (int64_t)lay.off[p] * d * e->elem, d, e->out_tab[p],
                l.logits + (int64_t)lay.off[p] * e->V, e->V, ActOut{}, n_p, rb::EPI_STORE, s, consumer(lay.off[p])));
  }
  return 0;
}

}  // namespace

extern "C" {

int rb200_engine_encode(rb200_engine* e, const int64_t* ids, const int64_t* mask, int batch, int S, int num_beams,
                        void* stream) {
  RB_REQUIRE(e && ids && mask, "null argument");
  if (!e->finalized) return rb::fail(RB200_ERR_STATE, "weights not finalized: call rb200_engine_finalize_weights");
  RB_REQUIRE(batch >= 1 && batch <= e->cfg.max_batch, "batch %d outside [1, %d]", batch, e->cfg.max_batch);
  RB_REQUIRE(S >= 1 && S <= e->cfg.max_src_len, "source length %d outside [1, %d]", S, e->cfg.max_src_len);
  RB_REQUIRE(num_beams >= 1 && num_beams <= e->cfg.max_beams, "num_beams %d outside [1, %d]", num_beams,
             e->cfg.max_beams);
  cudaStream_t s = (cudaStream_t)stream;
  RB_CUDA(cudaSetDevice(e->cfg.device));
  RB_TRY(build_bias_tables(e, S, s));
  return lane_encode(e, e->lanes[0], ids, mask, batch, S, num_beams, s);
}

int rb200_engine_encoder_states(const rb200_engine* e, const float** states) {
  RB_REQUIRE(e && states, "null argument");
  *states = e->lanes[0].enc_out;
  return 0;
}

int rb200_engine_decode_step(rb200_engine* e, const rb200_beam* beam, int t, float* logits, void* stream) {
  RB_REQUIRE(e && beam && logits, "null argument");
  Lane& l = e->lanes[0];
  if (l.B < 1) return rb::fail(RB200_ERR_STATE, "rb200_engine_encode has not been called");
  RB_REQUIRE(t >= 0 && t < e->Lmodel, "position %d outside [0, %d)", t, e->Lmodel);
  RB_REQUIRE(beam->step == t, "beam state is at step %d, decoder asked for position %d", beam->step, t);
  RB_REQUIRE(beam->nb == l.nb && beam->batch == l.B, "beam state shape differs from the encoded batch");
  RB_REQUIRE(beam->L == e->Lmodel, "beam state L=%d differs from the model's docid_len=%d", beam->L, e->Lmodel);
  return lane_decode_step(e, l, beam, t, logits, (cudaStream_t)stream);
}

int rb200_engine_beam(rb200_engine* e, rb200_beam** beam) {
  RB_REQUIRE(e && beam, "null argument");
  *beam = e->lanes[0].beam;
  return 0;
}

int64_t rb200_engine_last_launch_count(const rb200_engine* e) { return e ? e->launches : 0; }

int rb200_engine_last_tail_step(const rb200_engine* e) { return e ? e->last_tail_from : -1; }

int rb200_engine_next_input(const rb200_engine* e, float** next_x) {
  RB_REQUIRE(e && next_x, "null argument");
  *next_x = e->lanes[0].x_full;
  return 0;
}

int rb200_engine_resize(rb200_engine* e, int max_batch, int max_beams, int max_src_len) {
  RB_REQUIRE(e, "null argument");
  RB_REQUIRE(max_batch >= 1 && max_beams >= 1 && max_src_len >= 1, "max_batch, max_beams, max_src_len must be >= 1");
  if (max_batch == e->ws_batch && max_beams == e->ws_beams && max_src_len == e->ws_src) return 0;
  RB_CUDA(cudaSetDevice(e->cfg.device));
  RB_CUDA(cudaDeviceSynchronize());
  free_workspace(e);
  const int st = alloc_workspace(e, max_batch, max_beams, max_src_len);
  if (st != 0) {                      // keep the handle usable: nothing is allocated, the next resize may succeed
    free_workspace(e);
    e->ws_batch = e->ws_beams = e->ws_src = 0;
    e->cfg.max_batch = e->cfg.max_beams = e->cfg.max_src_len = 0;
  }
  return st;
}

int rb200_engine_last_freeze_histogram(const rb200_engine* e, int32_t* frozen_at, int n, int64_t* tail_rows) {
  RB_REQUIRE(e && frozen_at && n >= 1, "null argument");
  for (int t = 0; t < n; ++t) frozen_at[t] = t <= RB_TAIL_MAX_L ? e->frozen_at[t] : 0;
  if (tail_rows) *tail_rows = e->last_tail_rows;
  return 0;
}

int rb200_engine_search(rb200_engine* e, const rb200_trie* trie, const int64_t* ids, const int64_t* mask, int batch,
                        int S, int num_beams, int max_new_tokens, int num_return, int apply_log_softmax,
                        int64_t* sequences, float* scores, int32_t* leaf, void* stream) {
  RB_REQUIRE(e && trie && ids && mask && sequences && scores, "null argument");
  if (!e->finalized) return rb::fail(RB200_ERR_STATE, "weights not finalized: call rb200_engine_finalize_weights");
  RB_REQUIRE(batch >= 1 && batch <= e->cfg.max_batch, "batch %d outside [1, %d]", batch, e->cfg.max_batch);
  RB_REQUIRE(S >= 1 && S <= e->cfg.max_src_len, "source length %d outside [1, %d]", S, e->cfg.max_src_len);
  RB_REQUIRE(num_beams >= 1 && num_beams <= e->cfg.max_beams, "num_beams %d outside [1, %d]", num_beams,
             e->cfg.max_beams);
  RB_REQUIRE(max_new_tokens >= 1 && max_new_tokens <= e->Lmodel && max_new_tokens <= trie->L,
             "max_new_tokens=%d outside [1, min(model %d, trie %d)]", max_new_tokens, e->Lmodel, trie->L);
  RB_REQUIRE(num_return >= 1 && num_return <= num_beams,
             "`num_return_sequences` has to be smaller or equal to `num_beams`.");
  RB_REQUIRE(trie->V == e->V, "trie V=%d but the model's decoder_vocab_size is %d", trie->V, e->V);
  cudaStream_t s0 = (cudaStream_t)stream;
  RB_CUDA(cudaSetDevice(e->cfg.device));
  const int64_t launches0 = rb::launch_count();
  RB_CUDA(cudaMemsetAsync(e->overflow, 0, 4, s0));
  RB_TRY(build_bias_tables(e, S, s0));
  Lane& l = e->lanes[0];
  rb200_beam* beam = l.beam;
  RB_TRY(lane_encode(e, l, ids, mask, batch, S, num_beams, s0));
  RB_TRY(rb200_beam_reset_beams(beam, trie, batch, num_beams, s0));
  // Forced tail, per query: after every beam step the queries whose beams all sit on a single trie leaf are frozen
  // (their remaining tokens are determined by the trie) and leave the step loop; the loop goes on with the others as
  // a compacted batch. The host reads two counters back per step (stepping / frozen) to size the next launches.
  const int P = max_new_tokens;
  const bool try_tail = e->tail && P <= RB_TAIL_MAX_L;
  e->last_tail_from = -1;
  e->last_tail_rows = 0;
  for (int t = 0; t <= RB_TAIL_MAX_L; ++t) e->frozen_at[t] = 0;
  for (int t = 0; t < P && beam->n_active > 0; ++t) {
    const bool more = t + 1 < P;
    RB_TRY(lane_decode_step(e, l, beam, t, l.logits, s0));
    const int allow = try_tail && more;
    RB_TRY(rb::beam_step(beam, trie, l.logits, t == 0 ? 1 : num_beams, apply_log_softmax,
                         more ? e->in_tab[t] : nullptr, more ? l.x_full : nullptr, e->d, allow, s0));
    if (allow) {
      RB_TRY(rb::beam_compact(beam, s0));
      RB_CUDA(cudaMemcpyAsync(e->flag_host, beam->counts, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, s0));
      RB_CUDA(cudaStreamSynchronize(s0));
      const int n_act = e->flag_host[0], n_fz = e->flag_host[1];
      e->frozen_at[t + 1] = n_fz - beam->n_frozen;
      if (n_fz > beam->n_frozen && e->last_tail_from < 0) e->last_tail_from = t + 1;
      beam->n_active = n_act;
      beam->n_frozen = n_fz;
      beam->compacted = n_fz > 0;
    }
  }
  if (beam->n_frozen > 0) {
    rb::TailLayout lay;
    lay.P = P;
    int64_t off = 0, rows = 0;
    for (int p = 0; p <= RB_TAIL_MAX_L; ++p) {
      lay.off[p] = (int)off;
      if (p < P) {
        rows += (int64_t)e->frozen_at[p] * num_beams;     // block p: the beams of the queries frozen at or before p
        off += rows;
      }
    }
    RB_REQUIRE(off <= 0x7fffffff, "forced tail of %lld rows", (long long)off);
    e->last_tail_rows = off;
    RB_TRY(lane_tail(e, l, trie, lay, nullptr, s0));
    RB_TRY(rb::launch_tail_finish(beam, trie, lay, l.logits, apply_log_softmax, s0));
  }
  beam->step = P;
  beam->n_active = batch; beam->n_frozen = 0; beam->compacted = false;   // the final state is whole in the current half
  RB_TRY(rb200_beam_finalize(beam, trie, num_return, 1.0, sequences, scores, leaf, s0));
  if (rb::prec_is_fp16(e->mode)) {
    const int n = batch * num_return;
    poison_on_overflow_kernel<<<rb::ceil_div(n, 256), 256, 0, s0>>>(scores, n, e->overflow);
    RB_CUDA(cudaGetLastError());
    rb::launch_count()++;
  }
  e->launches = rb::launch_count() - launches0;
  return 0;
}

int rb200_engine_search_host(rb200_engine* e, const rb200_trie* trie, const int64_t* ids_host,
                             const int64_t* mask_host, int batch, int S, int num_beams, int max_new_tokens,
                             int num_return, int apply_log_softmax, int64_t* sequences_host, float* scores_host,
                             int32_t* leaf_host, void* stream) {
  RB_REQUIRE(e && ids_host && mask_host && sequences_host && scores_host, "null argument");
  RB_REQUIRE(batch >= 1 && batch <= e->cfg.max_batch, "batch %d outside [1, %d]", batch, e->cfg.max_batch);
  RB_REQUIRE(S >= 1 && S <= e->cfg.max_src_len, "source length %d outside [1, %d]", S, e->cfg.max_src_len);
  cudaStream_t s = (cudaStream_t)stream;
  RB_CUDA(cudaSetDevice(e->cfg.device));
  const size_t nin = (size_t)batch * S * 8;
  RB_CUDA(cudaMemcpyAsync(e->ids_dev, ids_host, nin, cudaMemcpyHostToDevice, s));
  RB_CUDA(cudaMemcpyAsync(e->mask_dev, mask_host, nin, cudaMemcpyHostToDevice, s));
  RB_TRY(rb200_engine_search(e, trie, e->ids_dev, e->mask_dev, batch, S, num_beams, max_new_tokens, num_return,
                             apply_log_softmax, e->seq_dev, e->score_dev, e->leaf_dev, stream));
  const size_t n = (size_t)batch * num_return;
  RB_CUDA(cudaMemcpyAsync(sequences_host, e->seq_dev, n * (max_new_tokens + 1) * 8, cudaMemcpyDeviceToHost, s));
  RB_CUDA(cudaMemcpyAsync(scores_host, e->score_dev, n * 4, cudaMemcpyDeviceToHost, s));
  if (leaf_host) RB_CUDA(cudaMemcpyAsync(leaf_host, e->leaf_dev, n * 2 * 4, cudaMemcpyDeviceToHost, s));
  RB_CUDA(cudaStreamSynchronize(s));
  return 0;
}

// Teacher-forced decoder pass over given DocID tokens (no beam search): what the reference's model forward computes
// when it is handed decoder_input_ids (t5_generative_retriever.py:295-450), and what rerank_forward sums (:794-798).
int rb200_engine_forward(rb200_engine* e, const int64_t* ids, const int64_t* mask, int batch, int S, int rows_per_query,
                         const int32_t* tokens, int T, float* logits_out, float* hidden_out, float* scores_out,
                         void* stream) {
  RB_REQUIRE(e && ids && mask && tokens, "null argument");
  if (!e->finalized) return rb::fail(RB200_ERR_STATE, "weights not finalized: call rb200_engine_finalize_weights");
  RB_REQUIRE(batch >= 1 && batch <= e->cfg.max_batch, "batch %d outside [1, %d]", batch, e->cfg.max_batch);
  RB_REQUIRE(S >= 1 && S <= e->cfg.max_src_len, "source length %d outside [1, %d]", S, e->cfg.max_src_len);
  RB_REQUIRE(rows_per_query >= 1 && rows_per_query <= e->cfg.max_beams, "rows_per_query %d outside [1, %d]",
             rows_per_query, e->cfg.max_beams);
  RB_REQUIRE(T >= 1 && T <= e->Lmodel && T <= RB_TAIL_MAX_L, "T=%d outside [1, min(model %d, %d)]", T, e->Lmodel,
             RB_TAIL_MAX_L);
  cudaStream_t s0 = (cudaStream_t)stream;
  RB_CUDA(cudaSetDevice(e->cfg.device));
  const int64_t launches0 = rb::launch_count();
  RB_CUDA(cudaMemsetAsync(e->overflow, 0, 4, s0));
  RB_TRY(build_bias_tables(e, S, s0));
  Lane& l = e->lanes[0];
  RB_TRY(lane_encode(e, l, ids, mask, batch, S, rows_per_query, s0));
  RB_TRY(rb::beam_force_tokens(l.beam, batch, rows_per_query, T, tokens, s0));
  const int R = batch * rows_per_query;
  rb::TailLayout lay;
  lay.P = T;
  for (int p = 0; p <= RB_TAIL_MAX_L; ++p) lay.off[p] = (p < T ? p : T) * R;
  RB_TRY(lane_tail(e, l, nullptr, lay, hidden_out, s0));
  const size_t nlog = (size_t)T * R * e->V;
  if (scores_out) RB_TRY(rb::launch_forced_scores(l.beam, lay, l.logits, scores_out, s0));
  if (logits_out) RB_CUDA(cudaMemcpyAsync(logits_out, l.logits, nlog * 4, cudaMemcpyDeviceToDevice, s0));
  if (rb::prec_is_fp16(e->mode)) {    // an activation left the fp16 range: unmistakably invalid results
    if (scores_out) poison_on_overflow_kernel<<<rb::ceil_div(R, 256), 256, 0, s0>>>(scores_out, R, e->overflow);
    if (logits_out) poison_on_overflow_kernel<<<1, 32, 0, s0>>>(logits_out, 1, e->overflow);
    RB_CUDA(cudaGetLastError());
  }
  l.beam->n_active = 0; l.beam->n_frozen = 0; l.beam->batch = 0;   // the beam state holds no search
  e->launches = rb::launch_count() - launches0;
  return 0;
}

int rb200_engine_set_profiling(rb200_engine* e, int on) {
  RB_REQUIRE(e, "null argument");
  e->profiling = on != 0;
  e->ev_used = 0;
  e->prof_flops = 0.0;
  e->prof_bytes = 0.0;
  return 0;
}

int rb200_engine_get_profile_bytes(const rb200_engine* e, double* gemm_bytes) {
  RB_REQUIRE(e && gemm_bytes, "null argument");
  *gemm_bytes = e->prof_bytes;
  return 0;
}

int rb200_engine_get_profile(rb200_engine* e, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches) {
  RB_REQUIRE(e && gemm_ms && gemm_flops && gemm_launches, "null argument");
  double ms = 0.0;
  for (size_t i = 0; i + 1 < e->ev_used; i += 2) {
    RB_CUDA(cudaEventSynchronize(e->events[i + 1]));
    float t = 0.f;
    RB_CUDA(cudaEventElapsedTime(&t, e->events[i], e->events[i + 1]));
    ms += t;
  }
  *gemm_ms = ms;
  *gemm_flops = e->prof_flops;
  *gemm_launches = (int64_t)(e->ev_used / 2);
  return 0;
}

int rb200_gemm(int precision, const float* A, const float* W, float* C, int64_t M, int64_t N, int64_t K,
               int accumulate, int relu, void* stream) {
  RB_REQUIRE(A && W && C, "null argument");
  RB_REQUIRE(precision >= 0 && precision <= 5, "unknown precision %d", precision);
  RB_REQUIRE(!(accumulate && relu), "accumulate and relu are exclusive");
  cudaStream_t s = (cudaStream_t)stream;
  const int planes = rb::prec_planes(precision), elem = rb::prec_elem_bytes(precision);
  void *Ap = nullptr, *Wp = nullptr, *Rp = nullptr;
  RB_CUDA(cudaMalloc(&Ap, (size_t)planes * M * K * elem));
  RB_CUDA(cudaMalloc(&Wp, (size_t)planes * N * K * elem));
  const float wscale = rb::prec_is_fp16(precision) ? rb::kFp16WeightScale : 1.0f;
  int st = rb::launch_pack_planes(A, Ap, M * K, M * K, precision, 1.0f, nullptr, s);
  if (st == 0) st = rb::launch_pack_planes(W, Wp, N * K, N * K, precision, wscale, nullptr, s);
  GemmArgs g;
  g.mode = precision; g.A = Ap; g.a_plane = M * K; g.W = Wp; g.w_plane = N * K;
  g.C = C; g.ldc = N; g.M = M; g.N = N; g.K = K;
  g.epilogue = accumulate ? rb::EPI_RESIDUAL : rb::EPI_STORE;
  g.out_scale = 1.0f / wscale;
  g.act = ActOut{};
  if (relu) {   // ReLU epilogue writes planes; unpack plane 0 (+ plane 1) back to fp32 for the caller
    if (cudaMalloc(&Rp, (size_t)planes * M * N * elem) != cudaSuccess) st = RB200_ERR_CUDA;
    g.epilogue = rb::EPI_RELU_ACT;
    g.act = ActOut{Rp, M * N, precision};
  }
  if (st == 0) st = rb::launch_gemm(g, s);
  if (st == 0 && relu) {
    std::vector<char> host((size_t)planes * M * N * elem);
    cudaStreamSynchronize(s);
    cudaMemcpy(host.data(), Rp, host.size(), cudaMemcpyDeviceToHost);
    std::vector<float> out((size_t)M * N);
    for (int64_t i = 0; i < M * N; ++i) {
      float v = 0.f;
      for (int p = 0; p < planes; ++p) {
        if (elem == 4) v += reinterpret_cast<float*>(host.data())[p * M * N + i];
        else {
          const uint16_t h16 = reinterpret_cast<uint16_t*>(host.data())[p * M * N + i];
          float f;
          if (rb::prec_is_fp16(precision)) {
            f = half_bits_to_float(h16);
          } else {
            const uint32_t b = (uint32_t)h16 << 16;
            memcpy(&f, &b, 4);
          }
          v += f;
        }
      }
      out[i] = v;
    }
    cudaMemcpy(C, out.data(), out.size() * 4, cudaMemcpyHostToDevice);
  }
  cudaStreamSynchronize(s);
  cudaFree(Ap); cudaFree(Wp); cudaFree(Rp);
  if (st == 0) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return rb::fail(RB200_ERR_CUDA, "rb200_gemm: %s", cudaGetErrorString(err));
  }
  return st;
}

int rb200_gemm_bench(int precision, int64_t M, int64_t N, int64_t K, int epilogue, int iters, int rotate_mb,
                     double* avg_us, void* stream) {
  RB_REQUIRE(avg_us && iters >= 1 && M >= 1 && N >= 1 && K >= 1, "bad argument");
  RB_REQUIRE(precision >= 0 && precision <= 5 && epilogue >= 0 && epilogue <= 2, "unknown precision / epilogue");
  cudaStream_t s = (cudaStream_t)stream;
  const int planes = rb::prec_planes(precision), elem = rb::prec_elem_bytes(precision);
  const size_t a_bytes = (size_t)planes * M * K * elem, w_bytes = (size_t)planes * N * K * elem;
  const size_t c_bytes = (size_t)M * N * 4, r_bytes = (size_t)planes * M * N * elem;
  const size_t per = a_bytes + w_bytes + c_bytes + r_bytes;
  int nbuf = (int)(((size_t)rotate_mb << 20) / per) + 1;
  if (nbuf > 64) nbuf = 64;
  std::vector<char*> bufs(nbuf, nullptr);
  int st = 0;
  for (int i = 0; i < nbuf && st == 0; ++i) {
    if (cudaMalloc((void**)&bufs[i], per + 4096) != cudaSuccess) st = rb::fail(RB200_ERR_NOMEM, "cudaMalloc failed");
    else cudaMemsetAsync(bufs[i], 0, per + 4096, s);   // zeros: valid in every plane format
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  auto run = [&](int i) {
    char* b = bufs[i % nbuf];
    GemmArgs g;
    g.mode = precision; g.A = b; g.a_plane = M * K; g.W = b + a_bytes; g.w_plane = N * K;
    g.C = reinterpret_cast<float*>(b + a_bytes + w_bytes); g.ldc = N; g.M = M; g.N = N; g.K = K;
    g.epilogue = epilogue; g.out_scale = 1.0f;
    g.act = epilogue == rb::EPI_RELU_ACT ? ActOut{b + a_bytes + w_bytes + c_bytes, M * N, precision, nullptr} : ActOut{};
    return rb::launch_gemm(g, s);
  };
  for (int i = 0; i < nbuf + 3 && st == 0; ++i) st = run(i);
  if (st == 0) {
    cudaEventRecord(e0, s);
    for (int i = 0; i < iters && st == 0; ++i) st = run(i);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    *avg_us = (double)ms * 1e3 / iters;
    if (getenv("RB200_GEMM_TRACE")) {   // one more launch with %globaltimer stamps per CTA; printed to stderr
      unsigned long long* tr = nullptr;
      const int slots = 160 * 8;
      cudaMalloc((void**)&tr, slots * 8);
      cudaMemsetAsync(tr, 0, slots * 8, s);
      rb::set_gemm_trace(tr);
      run(0);
      run(1);
      rb::set_gemm_trace(nullptr);
      cudaStreamSynchronize(s);
      std::vector<unsigned long long> h(slots);
      cudaMemcpy(h.data(), tr, slots * 8, cudaMemcpyDeviceToHost);
      cudaFree(tr);
      unsigned long long base = ~0ull;
      for (int c = 0; c < 160; ++c) if (h[c * 8] && h[c * 8] < base) base = h[c * 8];
      const char* names[7] = {"entry", "prologue done", "first stage full", "last MMA commit", "accumulator ready",
                              "epilogue done", "exit"};
      for (int k = 0; k < 7; ++k) {
        double sum = 0, mx = 0, mn = 1e30; int n = 0;
        for (int c = 0; c < 160; ++c) if (h[c * 8 + k]) {
          const double v = (double)(h[c * 8 + k] - base) / 1e3;
          sum += v; mx = v > mx ? v : mx; mn = v < mn ? v : mn; ++n;
        }
        if (n) fprintf(stderr, "  trace %-18s n=%3d  min %7.2f  avg %7.2f  max %7.2f us\n", names[k], n, mn, sum / n, mx);
      }
    }
  }
  cudaStreamSynchronize(s);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  for (char* b : bufs) cudaFree(b);
  if (st == 0) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return rb::fail(RB200_ERR_CUDA, "rb200_gemm_bench: %s", cudaGetErrorString(err));
  }
  return st;
}

}  // extern "C"
