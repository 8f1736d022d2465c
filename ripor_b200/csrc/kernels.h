// Launchers of the engine's non-tensor-core kernels and of the two GEMM back ends.
#pragma once
#include "act.cuh"
#include "beam.h"
#include "rb_common.h"

namespace rb {

enum GemmEpilogue {
  EPI_STORE = 0,       // C = acc                       (fp32)
  EPI_RESIDUAL = 1,    // C = C + acc                   (fp32 residual stream)
  EPI_RELU_ACT = 2,    // act_out = relu(acc)           (ActBuf for the next GEMM: DenseReluDense.wi)
  EPI_RESID_NORM = 3,  // C = C + acc, act_out = planes(C / r_prev), ss_out = per-row partial sums of C^2:
                       // the residual add with the FOLLOWING T5LayerNorm folded in (see NormFold below)
  EPI_PLANES = 4       // act_out = acc                 (operand planes, no ReLU: the forced tail's q | k | v and
                       // cross-attention q go straight to the tensor-core attention kernels as fp16 hi/lo planes)
};

// T5LayerNorm folded into the GEMMs around it (tensor-core modes). rmsnorm(x) * W^T = (1/r) * x * (W diag(ln))^T with
// r = sqrt(mean(x^2) + eps): the layer-norm weight is folded into the weight matrix when it is packed, the residual
// GEMM that produces x writes the operand planes of x / r_prev (r_prev = the row's r at the previous norm point, so
// the planes stay O(1) for the 16-bit formats) plus per-row partial sums of x^2 (one per 64 output columns, so the
// result does not depend on scheduling), and the consuming GEMM multiplies its accumulator rows by r_prev / r.
// That removes the 37 standalone RMSNorm launches of a decoder step.
struct NormFold {
  const float* ss_prev = nullptr;   // [M, np] partial sums of x^2 at the previous norm point
  const float* ss_cur = nullptr;    // [M, np] ... at this norm point (consumers only)
  float* ss_out = nullptr;          // [M, np] written by EPI_RESID_NORM
  int np = 0;                       // partial sums per row = d_model / 64
  float inv_d = 0.f, eps = 0.f;
  bool scaled = false;              // consumer: multiply accumulator rows by r_prev / r_cur
};

struct GemmArgs {
  int mode;              // rb200_precision
  const void* A;         // ActBuf planes [planes][a_rows_cap][K]
  int64_t a_plane;       // elements between A planes
  const void* W;         // packed weight planes [planes][N][K]
  int64_t w_plane;
  float* C;              // fp32 output, leading dimension ldc (EPI_STORE / EPI_RESIDUAL)
  int64_t ldc;
  ActOut act;            // EPI_RELU_ACT output (row length N)
  int64_t M, N, K;
  int epilogue;
  float out_scale = 1.0f;  // accumulator multiplier (undoes the power-of-two pre-scale of fp16 weight planes)
  NormFold nf;
};

// fp16 weight planes are stored multiplied by 2^8 so that the low plane stays in fp16's normal range
constexpr float kFp16WeightScale = 256.0f;

// x[row, :] = table[ids[row], :]            (encoder token embedding, shared.weight)
int launch_embed_rows(const float* table, const int64_t* ids, float* x, int64_t rows, int d, cudaStream_t s);
// x[row, :] = vec[:]                         (decoder start_token_embed)
int launch_broadcast_row(const float* vec, float* x, int64_t rows, int d, cudaStream_t s);
// T5LayerNorm: out = w * x * rsqrt(mean(x^2) + eps) * scale, written as an ActBuf
int launch_rmsnorm(const float* x, const float* w, ActOut out, int64_t rows, int d, float eps, float scale,
                   cudaStream_t s);
// first norm point of a chain under NormFold: out = planes(x / r), ss[0][row][:] = ss[1][row][:] = {sum x^2, 0, ...}
int launch_norm_init(const float* x, ActOut out, float* ss0, float* ss1, int np, int64_t rows, int d, float eps,
                     cudaStream_t s);
// same, fp32 output (encoder last_hidden_state)
int launch_rmsnorm_f32(const float* x, const float* w, float* out, int64_t rows, int d, float eps, cudaStream_t s,
                       float scale = 1.0f);

struct SelfAttnArgs {
  const float* qkv;       // [M, 3*inner] (q | k | v) of the current position
  float* cache_k;         // [L, row_cap, inner] of this layer
  float* cache_v;
  const int32_t* anc;     // [R, L] KV ancestry (beam state)
  const float* bias;      // [H, L] relative position bias by distance (t - p)
  int64_t row_cap;
  int M, H, L, t, rpq, nb;
  // compact rows (some queries have left the step loop): row m belongs to query qlist[m / nb]; the cache slots and
  // the ancestry table keep the ORIGINAL row ids. nullptr = identity.
  const int32_t* qlist = nullptr;
};
int launch_self_attn_decode(const SelfAttnArgs& a, ActOut ctx, cudaStream_t s);

struct CrossAttnArgs {
  const float* q;         // [M, q_ld] (q_ld = 0: rows of `inner` floats)
  const float* kv;        // cross K/V of all layers [B*S, ld]; this layer's K at k_off, V at v_off
  int64_t ld, k_off, v_off;
  const int64_t* mask;    // [B, S] encoder attention mask
  int M, H, S, rows_per_query;
  int64_t q_ld = 0;
  // forced tail: nblocks position blocks of block_rows rows each (row = block * block_rows + query * rpq + beam)
  int nblocks = 1;
  int64_t block_rows = 0;
  // encoder self-attention through the same kernel: the "beams" are the S query rows of the sequence and every score
  // gets the relative position bias rel_bias[h][(key - query) + S - 1]
  const float* rel_bias = nullptr;
  int rel_S = 0;            // the table was built for sequences of up to rel_S positions: [H, 2*rel_S-1]
  // compact query index -> query whose encoder K/V and mask to use (step loop: the stepping list; forced tail: the
  // freeze-order list). nullptr = identity.
  const int32_t* qmap = nullptr;
  // ragged forced tail: query slot fq runs positions qstart[qmap[fq]] (0 if qstart is null) .. lay.P-1 and its rows
  // are lay.off[p] + fq * rows_per_query + beam (nblocks / block_rows are then ignored; M = slots * rows_per_query)
  int ragged = 0;
  const int32_t* qstart = nullptr;
  TailLayout lay;
};
int launch_cross_attn_decode(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s);

// Self-attention of the forced tail: every frozen row r' (freeze order) runs its remaining positions t0..P-1 in one
// task. Rows are position-major and ragged (row = lay.off[p] + r'); positions < t0 come from the KV cache through the
// frozen ancestry table (indexed by the ORIGINAL row id), positions >= t0 from the rows of this pass. P <= 32.
struct TailAttnArgs {
  const float* qkv;       // [rows of the pass, 3*inner]
  const float* cache_k;   // this layer's cache [L, row_cap, inner]
  const float* cache_v;
  const int32_t* anc;     // [R, L] frozen ancestry
  const float* bias;      // [H, L]
  int64_t row_cap;
  int R;                  // frozen rows = frozen queries * nb
  int H, L, nb;
  const int32_t* fz_list; // frozen slot -> query
  const int32_t* qstart;  // per query: first position of the pass (nullptr: 0 for all, no cached prefix)
  TailLayout lay;
  // fp16x3 mode: q | k | v of the pass as fp16 hi/lo planes written by the EPI_PLANES GEMM epilogue (hi plane at
  // qkv_hi, lo plane qkv_plane elements further; same [rows, 3*inner] layout). nullptr: fp32 rows in `qkv`.
  const __half* qkv_hi = nullptr;
  int64_t qkv_plane = 0;
};
bool tail_self_attn_reads_planes(int mode);
int launch_self_attn_tail(const TailAttnArgs& a, ActOut ctx, cudaStream_t s);

// fp32 FFMA GEMM (RB200_PREC_FP32) and the tcgen05 GEMM family (all other modes)
int launch_gemm_simt(const GemmArgs& g, cudaStream_t s);
int launch_gemm_sm100(const GemmArgs& g, cudaStream_t s);
inline int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  return g.mode == RB200_PREC_FP32 ? launch_gemm_simt(g, s) : launch_gemm_sm100(g, s);
}

// fp32 [rows, cols] -> packed planes of `mode` (weights at load time, and rb200_gemm's operands)
// every element is multiplied by `scale` first; `overflow` (device int, may be null) is raised by fp16 planes
int launch_pack_planes(const float* src, void* dst, int64_t numel, int64_t plane, int mode, float scale,
                       int* overflow, cudaStream_t s);
// same with a per-column factor (row length k): dst = planes(src * scale * col_scale[col]) - layer norm folding
int launch_pack_planes_cols(const float* src, void* dst, int64_t numel, int64_t plane, int mode, float scale,
                            const float* col_scale, int64_t k, int* overflow, cudaStream_t s);

// HF T5 relative position bucket (host; float32 arithmetic like torch)
int relative_bucket(int rel, bool bidirectional, int num_buckets, int max_distance);

}  // namespace rb
