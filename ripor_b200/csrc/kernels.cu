// Non-tensor-core kernels of the T5 encoder / KV-cached decoder, plus the exact fp32 FFMA GEMM.
//
// Arithmetic follows HF T5 as the reference drives it (t5_pretrainer/modeling/t5_generative_retriever.py
// :358-366 encoder, :403-416 decoder): T5LayerNorm without mean subtraction, attention WITHOUT 1/sqrt(dk)
// scaling, additive relative position bias shared by all blocks, softmax in fp32, hard masks where HF adds
// a large negative number. Everything here is HBM/L2-bound gather + reduce work in fp32.
#include <cmath>

#include "kernels.h"

namespace rb {
namespace {

// ------------------------------------------------------------------------------------------------
// row gathers / broadcast
// ------------------------------------------------------------------------------------------------
__global__ void embed_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ ids,
                                  float* __restrict__ x, int64_t rows, int d4) {
  const int64_t row = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(table) + ids[row] * d4;
  float4* dst = reinterpret_cast<float4*>(x) + row * d4;
  for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = src[c];
}

__global__ void broadcast_row_kernel(const float* __restrict__ vec, float* __restrict__ x, int64_t rows, int d4) {
  const int64_t row = blockIdx.x;
  const float4* src = reinterpret_cast<const float4*>(vec);
  float4* dst = reinterpret_cast<float4*>(x) + row * d4;
  for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = src[c];
}

// ------------------------------------------------------------------------------------------------
// T5LayerNorm (one warp per row)
// ------------------------------------------------------------------------------------------------
template <bool F32OUT, int NV>   // NV float4 per lane held in registers (d <= 128 * NV); NV = 0: two-pass fallback
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, ActOut out,
                               float* __restrict__ out_f32, int64_t rows, int d, float eps, float scale) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_wait();
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  const float4* wr = reinterpret_cast<const float4*>(w);
  const int d4 = d >> 2;
  float4 v[NV > 0 ? NV : 1];
  float ss = 0.f;
  if (NV > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < d4 ? xr[c] : make_float4(0.f, 0.f, 0.f, 0.f);     // the row is read once and stays in registers
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) ss += v[i].x * v[i].x + v[i].y * v[i].y + v[i].z * v[i].z + v[i].w * v[i].w;
  } else {
    for (int c = lane; c < d4; c += 32) {
      const float4 t = xr[c];
      ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    }
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rs = 1.0f / sqrtf(ss / (float)d + eps);
  pdl_trigger();   // late: an early trigger parks the next kernel's CTAs on SMs this grid still needs
  auto emit = [&](int c, float4 t) {
    const float4 g = __ldg(wr + c);
    float4 y = make_float4(g.x * (t.x * rs), g.y * (t.y * rs), g.z * (t.z * rs), g.w * (t.w * rs));
    if (scale != 1.0f) { y.x *= scale; y.y *= scale; y.z *= scale; y.w *= scale; }
    if (F32OUT) reinterpret_cast<float4*>(out_f32 + row * d)[c] = y;
    else act_store4(out, row * d + (int64_t)c * 4, y);
  };
  if (NV > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i)
      if (lane + 32 * i < d4) emit(lane + 32 * i, v[i]);
  } else {
    for (int c = lane; c < d4; c += 32) emit(c, xr[c]);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 FFMA GEMM: C[M,N] = A[M,K] * W[N,K]^T, 128x128x16 tiles, 8x8 per thread. Exact fp32 products and
// fp32 accumulation: the reference-arithmetic mode and the yardstick the tensor-core modes are tested on.
// ------------------------------------------------------------------------------------------------
constexpr int SBM = 128, SBN = 128, SBK = 16;

__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                        float* __restrict__ C, int64_t ldc, ActOut act, int64_t M,
                                                        int64_t N, int64_t K, int epilogue) {
  __shared__ __align__(16) float As[SBK][SBM + 4];
  __shared__ __align__(16) float Ws[SBK][SBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * SBM, n0 = (int64_t)blockIdx.x * SBN;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  // each thread loads 2 float4 of A and 2 of W per k-tile: row = (tid*2+u)/4, k4 = (tid*2+u)%4
  for (int64_t k0 = 0; k0 < K; k0 += SBK) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int f = tid * 2 + u, row = f >> 2, k4 = (f & 3) * 4;
      float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vw = va;
      if (m0 + row < M && k0 + k4 < K) va = *reinterpret_cast<const float4*>(A + (m0 + row) * K + k0 + k4);
      if (n0 + row < N && k0 + k4 < K) vw = *reinterpret_cast<const float4*>(W + (n0 + row) * K + k0 + k4);
      As[k4 + 0][row] = va.x; As[k4 + 1][row] = va.y; As[k4 + 2][row] = va.z; As[k4 + 3][row] = va.w;
      Ws[k4 + 0][row] = vw.x; Ws[k4 + 1][row] = vw.y; Ws[k4 + 2][row] = vw.z; Ws[k4 + 3][row] = vw.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Ws[k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int64_t n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      const float v = acc[i][j];
      if (epilogue == EPI_STORE) C[m * ldc + n] = v;
      else if (epilogue == EPI_RESIDUAL) C[m * ldc + n] += v;
      else act_store(act, m * N + n, fmaxf(v, 0.f));
    }
  }
}

__global__ void pack_planes_kernel(const float* __restrict__ src, ActOut out, int64_t numel, float scale,
                                   const float* __restrict__ col_scale, int64_t k) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < numel) act_store(out, i, col_scale ? (src[i] * col_scale[i % k]) * scale : src[i] * scale);
}

// NormFold chain start: one warp per row; planes(x / r), and both parities of the partial-sum table
__global__ void norm_init_kernel(const float* __restrict__ x, ActOut out, float* __restrict__ ss0,
                                 float* __restrict__ ss1, int np, int64_t rows, int d, float eps) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_wait();
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  const int d4 = d >> 2;
  float ss = 0.f;
  for (int c = lane; c < d4; c += 32) {
    const float4 t = xr[c];
    ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rs = 1.0f / sqrtf(ss / (float)d + eps);
  pdl_trigger();
  for (int c = lane; c < d4; c += 32) {
    const float4 t = xr[c];
    act_store4(out, row * d + (int64_t)c * 4, make_float4(t.x * rs, t.y * rs, t.z * rs, t.w * rs));
  }
  for (int i = lane; i < np; i += 32) {
    const float v = i == 0 ? ss : 0.f;
    ss0[row * np + i] = v;
    ss1[row * np + i] = v;
  }
}

}  // namespace

int launch_embed_rows(const float* table, const int64_t* ids, float* x, int64_t rows, int d, cudaStream_t s) {
  RB_REQUIRE(d % 4 == 0, "d_model must be a multiple of 4");
  if (rows == 0) return 0;
  embed_rows_kernel<<<(unsigned)rows, 128, 0, s>>>(table, ids, x, rows, d / 4);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

int launch_broadcast_row(const float* vec, float* x, int64_t rows, int d, cudaStream_t s) {
  if (rows == 0) return 0;
  broadcast_row_kernel<<<(unsigned)rows, 128, 0, s>>>(vec, x, rows, d / 4);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

template <bool F32OUT>
static int launch_rmsnorm_any(const float* x, const float* w, ActOut out, float* out_f32, int64_t rows, int d, float eps,
                              float scale, cudaStream_t s) {
  if (rows == 0) return 0;
  const dim3 grid(ceil_div(rows, 8)), block(256);     // one warp per row, 8 rows per CTA
  cudaError_t err;
  if (d <= 128 * 6) err = launch_pdl(rmsnorm_kernel<F32OUT, 6>, grid, block, 0, s, x, w, out, out_f32, rows, d, eps, scale);
  else if (d <= 128 * 8) err = launch_pdl(rmsnorm_kernel<F32OUT, 8>, grid, block, 0, s, x, w, out, out_f32, rows, d, eps, scale);
  else err = launch_pdl(rmsnorm_kernel<F32OUT, 0>, grid, block, 0, s, x, w, out, out_f32, rows, d, eps, scale);
  RB_CUDA(err);
  rb::launch_count()++;
  return 0;
}

int launch_rmsnorm(const float* x, const float* w, ActOut out, int64_t rows, int d, float eps, float scale,
                   cudaStream_t s) {
  return launch_rmsnorm_any<false>(x, w, out, nullptr, rows, d, eps, scale, s);
}

int launch_rmsnorm_f32(const float* x, const float* w, float* out, int64_t rows, int d, float eps, cudaStream_t s,
                       float scale) {
  return launch_rmsnorm_any<true>(x, w, ActOut{nullptr, 0, 0}, out, rows, d, eps, scale, s);
}

// attn_warp.cu: the attention kernels (one warp per (row, head)); the encoder reuses the cross-attention kernel
bool launch_self_attn_warp(const SelfAttnArgs& a, ActOut ctx, cudaStream_t s, int* status);
bool launch_cross_attn_warp(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s, int* status);

int launch_self_attn_decode(const SelfAttnArgs& a, ActOut ctx, cudaStream_t s) {
  int st = 0;
  launch_self_attn_warp(a, ctx, s, &st);
  return st;
}

int launch_cross_attn_decode(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s) {
  RB_REQUIRE(a.M % a.rows_per_query == 0, "row count %d is not a multiple of rows_per_query %d", a.M,
             a.rows_per_query);
  int st = 0;
  launch_cross_attn_warp(a, ctx, s, &st);
  return st;
}

int launch_gemm_simt(const GemmArgs& g, cudaStream_t s) {
  RB_REQUIRE(g.K % 4 == 0, "K=%lld must be a multiple of 4", (long long)g.K);
  if (g.M == 0 || g.N == 0) return 0;
  dim3 grid(ceil_div(g.N, SBN), ceil_div(g.M, SBM));
  gemm_simt_kernel<<<grid, 256, 0, s>>>(static_cast<const float*>(g.A), static_cast<const float*>(g.W), g.C, g.ldc,
                                        g.act, g.M, g.N, g.K, g.epilogue);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

int launch_pack_planes_cols(const float* src, void* dst, int64_t numel, int64_t plane, int mode, float scale,
                            const float* col_scale, int64_t k, int* overflow, cudaStream_t s) {
  if (numel == 0) return 0;
  pack_planes_kernel<<<ceil_div(numel, 256), 256, 0, s>>>(src, ActOut{dst, plane, mode, overflow}, numel, scale,
                                                          col_scale, k);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

int launch_pack_planes(const float* src, void* dst, int64_t numel, int64_t plane, int mode, float scale,
                       int* overflow, cudaStream_t s) {
  return launch_pack_planes_cols(src, dst, numel, plane, mode, scale, nullptr, 1, overflow, s);
}

int launch_norm_init(const float* x, ActOut out, float* ss0, float* ss1, int np, int64_t rows, int d, float eps,
                     cudaStream_t s) {
  if (rows == 0) return 0;
  RB_CUDA(launch_pdl(norm_init_kernel, dim3(ceil_div(rows, 8)), dim3(256), 0, s, x, out, ss0, ss1, np, rows, d, eps));
  rb::launch_count()++;
  return 0;
}

int relative_bucket(int rel, bool bidirectional, int num_buckets, int max_distance) {
  // HF T5Attention._relative_position_bucket with torch's float32 arithmetic.
  int bucket = 0;
  if (bidirectional) {
    num_buckets /= 2;
    if (rel > 0) bucket += num_buckets;
    rel = rel < 0 ? -rel : rel;
  } else {
    rel = rel < 0 ? -rel : 0;
  }
  const int max_exact = num_buckets / 2;
  if (rel < max_exact) return bucket + rel;
  const float ratio = (float)rel / (float)max_exact;
  const float denom = (float)std::log((double)max_distance / (double)max_exact);
  float v = std::log(ratio) / denom * (float)(num_buckets - max_exact);
  int large = max_exact + (int)v;
  if (large > num_buckets - 1) large = num_buckets - 1;
  return bucket + large;
}

}  // namespace rb
