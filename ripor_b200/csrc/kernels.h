// Launchers of the engine's non-tensor-core kernels and of the two GEMM back ends.
#pragma once
#include "act.cuh"
#include "rb_common.h"

namespace rb {

enum GemmEpilogue {
  EPI_STORE = 0,     // C = acc                       (fp32)
  EPI_RESIDUAL = 1,  // C = C + acc                   (fp32 residual stream)
  EPI_RELU_ACT = 2   // act_out = relu(acc)           (ActBuf for the next GEMM: DenseReluDense.wi)
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
// same, fp32 output (encoder last_hidden_state)
int launch_rmsnorm_f32(const float* x, const float* w, float* out, int64_t rows, int d, float eps, cudaStream_t s);

struct SelfAttnArgs {
  const float* qkv;       // [M, 3*inner] (q | k | v) of the current position
  float* cache_k;         // [L, row_cap, inner] of this layer
  float* cache_v;
  const int32_t* anc;     // [R, L] KV ancestry (beam state)
  const float* bias;      // [H, L] relative position bias by distance (t - p)
  int64_t row_cap;
  int M, H, L, t, rpq, nb;
};
int launch_self_attn_decode(const SelfAttnArgs& a, ActOut ctx, cudaStream_t s);

struct CrossAttnArgs {
  const float* q;         // [M, inner]
  const float* kv;        // cross K/V of all layers [B*S, ld]; this layer's K at k_off, V at v_off
  int64_t ld, k_off, v_off;
  const int64_t* mask;    // [B, S] encoder attention mask
  int M, H, S, rows_per_query;
};
int launch_cross_attn_decode(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s);

struct EncAttnArgs {
  const float* qkv;       // [B*S, 3*inner]
  const int64_t* mask;    // [B, S]
  const float* bias;      // [H, 2*S-1] relative position bias by (key - query) + S - 1
  int B, S, H;
};
int launch_enc_attn(const EncAttnArgs& a, ActOut ctx, cudaStream_t s);

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

// HF T5 relative position bucket (host; float32 arithmetic like torch)
int relative_bucket(int rel, bool bidirectional, int num_buckets, int max_distance);

}  // namespace rb
