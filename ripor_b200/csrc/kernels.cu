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
template <bool F32OUT>
__global__ void rmsnorm_kernel(const float* __restrict__ x, const float* __restrict__ w, ActOut out,
                               float* __restrict__ out_f32, int64_t rows, int d, float eps, float scale) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_trigger();
  pdl_wait();
  if (row >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + row * d);
  const float4* wr = reinterpret_cast<const float4*>(w);
  const int d4 = d >> 2;
  float ss = 0.f;
  for (int c = lane; c < d4; c += 32) {
    const float4 v = xr[c];
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rs = 1.0f / sqrtf(ss / (float)d + eps);
  for (int c = lane; c < d4; c += 32) {
    const float4 v = xr[c];
    const float4 g = wr[c];
    float4 y = make_float4(g.x * (v.x * rs), g.y * (v.y * rs), g.z * (v.z * rs), g.w * (v.w * rs));
    if (scale != 1.0f) { y.x *= scale; y.y *= scale; y.z *= scale; y.w *= scale; }
    if (F32OUT) reinterpret_cast<float4*>(out_f32 + row * d)[c] = y;
    else act_store4(out, row * d + (int64_t)c * 4, y);
  }
}

// ------------------------------------------------------------------------------------------------
// row-wise attention: one CTA per query row, thread c owns float4 chunk c of the inner dimension, the
// 16 threads of a head reduce the q.k dot products with shuffles; scores live in shared memory.
// ------------------------------------------------------------------------------------------------
template <typename KPtr, typename VPtr, typename Valid, typename Bias>
__device__ __forceinline__ float4 attn_core(float4 q4, int h, int c, bool active, int H, int P, float* sc,
                                            KPtr kptr, VPtr vptr, Valid valid, Bias bias) {
  const int lane16 = threadIdx.x & 15;
#pragma unroll 4
  for (int p = 0; p < P; ++p) {
    float part = 0.f;
    const bool ok = valid(p);
    if (ok && active) {
      const float4 k4 = __ldg(kptr(p) + c);
      part = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
    }
    part += __shfl_xor_sync(0xffffffffu, part, 8);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    if (active && lane16 == 0) sc[h * P + p] = ok ? part + bias(h, p) : -INFINITY;
  }
  __syncthreads();
  {   // softmax over the P positions of head h by its 16 threads (inactive lanes shadow head 0, read-only)
    const int hs = active ? h : 0;
    float m = -INFINITY;
    for (int p = lane16; p < P; p += 16) m = fmaxf(m, sc[hs * P + p]);
    for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o, 16));
    float sum = 0.f;
    for (int p = lane16; p < P; p += 16) sum += expf(sc[hs * P + p] - m);
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 16);
    __syncthreads();
    if (active)
      for (int p = lane16; p < P; p += 16) sc[h * P + p] = expf(sc[h * P + p] - m) / sum;
  }
  __syncthreads();
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
#pragma unroll 4
    for (int p = 0; p < P; ++p) {
      if (!valid(p)) continue;
      const float pr = sc[h * P + p];
      const float4 v4 = __ldg(vptr(p) + c);
      o.x += pr * v4.x; o.y += pr * v4.y; o.z += pr * v4.z; o.w += pr * v4.w;
    }
  }
  return o;
}

__global__ void self_attn_decode_kernel(SelfAttnArgs a, ActOut ctx) {
  extern __shared__ float smem[];
  const int inner = a.H * 64, c4n = inner >> 2;
  const int m = blockIdx.x, c = threadIdx.x;
  const bool active = c < c4n;
  const int h = c >> 4;
  const int P = a.t + 1;
  int* anc_s = reinterpret_cast<int*>(smem);          // [L]
  float* sc = smem + a.L;                              // [H, P]
  const int arow = (a.rpq == 1) ? m * a.nb : m;
  pdl_trigger();
  pdl_wait();
  for (int p = threadIdx.x; p < P; p += blockDim.x) anc_s[p] = a.anc[(int64_t)arow * a.L + p];
  float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    const float4* row = reinterpret_cast<const float4*>(a.qkv + (int64_t)m * 3 * inner);
    q4 = row[c];
    // this position's K/V go to the cache slot (t, m); the same thread reads them back below
    float4* ck = reinterpret_cast<float4*>(a.cache_k + ((int64_t)a.t * a.row_cap + m) * inner);
    float4* cv = reinterpret_cast<float4*>(a.cache_v + ((int64_t)a.t * a.row_cap + m) * inner);
    ck[c] = row[c4n + c];
    cv[c] = row[2 * c4n + c];
  }
  __syncthreads();
  const float* ck = a.cache_k;
  const float* cv = a.cache_v;
  const int64_t rc = a.row_cap;
  const int t = a.t, L = a.L;
  const float* bias = a.bias;
  // K/V of position t were just written by this thread: read with plain loads (not the read-only path)
  auto kp = [&](int p) { return reinterpret_cast<const float4*>(ck + ((int64_t)p * rc + anc_s[p]) * inner); };
  auto vp = [&](int p) { return reinterpret_cast<const float4*>(cv + ((int64_t)p * rc + anc_s[p]) * inner); };
  const int lane16 = threadIdx.x & 15;
  // phase 1 inlined (cannot use __ldg on the freshly written slot)
  for (int p = 0; p < P; ++p) {
    float part = 0.f;
    if (active) {
      const float4 k4 = kp(p)[c];
      part = q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
    }
    part += __shfl_xor_sync(0xffffffffu, part, 8);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    if (active && lane16 == 0) sc[h * P + p] = part + bias[h * L + (t - p)];
  }
  __syncthreads();
  {
    const int hs = active ? h : 0;
    float mx = -INFINITY;
    for (int p = lane16; p < P; p += 16) mx = fmaxf(mx, sc[hs * P + p]);
    for (int o = 8; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o, 16));
    float sum = 0.f;
    for (int p = lane16; p < P; p += 16) sum += expf(sc[hs * P + p] - mx);
    for (int o = 8; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o, 16);
    __syncthreads();
    if (active)
      for (int p = lane16; p < P; p += 16) sc[h * P + p] = expf(sc[h * P + p] - mx) / sum;
  }
  __syncthreads();
  if (active) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < P; ++p) {
      const float pr = sc[h * P + p];
      const float4 v4 = vp(p)[c];
      o.x += pr * v4.x; o.y += pr * v4.y; o.z += pr * v4.z; o.w += pr * v4.w;
    }
    act_store4(ctx, (int64_t)m * inner + (int64_t)c * 4, o);
  }
}

// Cross-attention for the decoder step. One CTA per (query, chunk of <= NBC beams, group of HG heads); warp = head.
// The query's K/V rows are identical for all of its beams (the reference expands them x num_beams,
// generation.py:231-233, we never do): a 32-position K (then V) chunk is staged ONCE in shared memory and
// reused by every beam of the chunk. Phase 1: lane = key position, a full 64-dim dot product per (beam, position)
// from shared memory (no shuffle reductions: the first version of this kernel spent its time issuing 4
// shuffles + adds per 4-element partial dot). Phase 2: lane = a pair of output dims.
constexpr int NBC = 10;          // beams per CTA
constexpr int XHG = 4;           // heads per CTA
constexpr int XLD = XHG * 64 + 4;  // padded row of the staged K/V chunk (floats): conflict-free 128-bit rows

__global__ void __launch_bounds__(XHG * 32) cross_attn_decode_kernel(CrossAttnArgs a, ActOut ctx) {
  extern __shared__ __align__(16) float smem[];
  const int inner = a.H * 64;
  const int b = blockIdx.x, hg0 = blockIdx.z * XHG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = hg0 + warp;                      // this warp's head
  const bool head_ok = h < a.H;
  const int rpq = a.rows_per_query, S = a.S;
  const int i0 = blockIdx.y * NBC;
  const int nact = min(NBC, rpq - i0);
  const int64_t row0 = (int64_t)b * rpq + i0;
  float* k_s = smem;                             // [32][XLD] staged K chunk
  float* v_s = k_s + 32 * XLD;                   // [32][XLD] staged V chunk (prefetched while phase 1 runs)
  float* q_s = v_s + 32 * XLD;                   // [NBC][XHG*64]
  float* sc_s = q_s + NBC * XHG * 64;            // [NBC][XHG][S]
  pdl_trigger();
  pdl_wait();
  const int cols = min(XHG, a.H - hg0) * 64;     // valid floats per staged row
  for (int e = threadIdx.x; e < NBC * XHG * 16; e += blockDim.x) {
    const int i = e / (XHG * 16), c4 = e - i * (XHG * 16);
    const bool ok = i < nact && c4 * 4 < cols;
    const float* src = a.q + (row0 + (ok ? i : 0)) * inner + hg0 * 64 + (ok ? c4 * 4 : 0);
    const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(q_s + i * XHG * 64 + c4 * 4);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d32), "l"(src), "r"(ok ? 16 : 0) : "memory");
  }   // joins the commit group of the first K chunk below
  const float* base = a.kv + (int64_t)b * S * a.ld + hg0 * 64;
  const int64_t* mk = a.mask + (int64_t)b * S;
  // 32 positions x cols floats with cp.async (16 B, zero-filled for masked / out-of-range rows): the loads of a
  // chunk are all in flight at once instead of one L2 round trip per loop iteration
  auto stage = [&](float* dst, int p0, int64_t off) {
    for (int e = threadIdx.x; e < 32 * XHG * 16; e += blockDim.x) {
      const int r = e / (XHG * 16), c4 = e - r * (XHG * 16);
      const int p = p0 + r;
      const bool ok = p < S && c4 * 4 < cols && mk[p < S ? p : 0] != 0;
      const float* src = base + (int64_t)(ok ? p : 0) * a.ld + off + (ok ? c4 * 4 : 0);
      const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst + r * XLD + c4 * 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d32), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // ---- phase 1: scores[i][head][p] = q_i . k_p -------------------------------------------------------
  stage(k_s, 0, a.k_off);
  stage(v_s, 0, a.v_off);                        // V chunk 0 arrives while phase 1 computes
  for (int p0 = 0; p0 < S; p0 += 32) {
    if (p0 > 0) {
      __syncthreads();
      stage(k_s, p0, a.k_off);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    }
    __syncthreads();
    float acc[NBC];
#pragma unroll
    for (int i = 0; i < NBC; ++i) acc[i] = 0.f;
    const float* krow = k_s + lane * XLD + warp * 64;
    const float* qh = q_s + warp * 64;
#pragma unroll 4
    for (int d4 = 0; d4 < 16; ++d4) {
      const float4 k4 = *reinterpret_cast<const float4*>(krow + d4 * 4);
#pragma unroll
      for (int i = 0; i < NBC; ++i) {
        const float4 q4 = *reinterpret_cast<const float4*>(qh + i * XHG * 64 + d4 * 4);
        acc[i] += q4.x * k4.x + q4.y * k4.y + q4.z * k4.z + q4.w * k4.w;
      }
    }
    const int p = p0 + lane;
    if (p < S) {
      const bool ok = mk[p] != 0;
#pragma unroll
      for (int i = 0; i < NBC; ++i)
        if (i < nact) sc_s[(i * XHG + warp) * S + p] = ok ? acc[i] : -INFINITY;
    }
  }
  __syncthreads();
  // ---- softmax over the S positions of every (beam, head) row: one warp per head -----------------------
  for (int i = 0; i < nact; ++i) {
    float* sc = sc_s + (i * XHG + warp) * S;
    float m = -INFINITY;
    for (int p = lane; p < S; p += 32) m = fmaxf(m, sc[p]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int p = lane; p < S; p += 32) sum += expf(sc[p] - m);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    for (int p = lane; p < S; p += 32) sc[p] = expf(sc[p] - m) / sum;
  }
  // ---- phase 2: out[i][head][2*lane .. +1] = sum_p prob[i][p] * v_p ------------------------------------
  float2 o2[NBC];
#pragma unroll
  for (int i = 0; i < NBC; ++i) o2[i] = make_float2(0.f, 0.f);
  for (int p0 = 0; p0 < S; p0 += 32) {
    if (p0 > 0) {
      __syncthreads();
      stage(v_s, p0, a.v_off);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int np = min(32, S - p0);
    for (int r = 0; r < np; ++r) {
      const float2 v2 = *reinterpret_cast<const float2*>(v_s + r * XLD + warp * 64 + lane * 2);
#pragma unroll
      for (int i = 0; i < NBC; ++i) {
        const float pr = (i < nact) ? sc_s[(i * XHG + warp) * S + p0 + r] : 0.f;   // masked keys: exactly 0
        o2[i].x += pr * v2.x;
        o2[i].y += pr * v2.y;
      }
    }
  }
  if (head_ok) {
#pragma unroll
    for (int i = 0; i < NBC; ++i)
      if (i < nact) {
        const int64_t idx = (row0 + i) * inner + h * 64 + lane * 2;
        act_store(ctx, idx, o2[i].x);
        act_store(ctx, idx + 1, o2[i].y);
      }
  }
}

__global__ void enc_attn_kernel(EncAttnArgs a, ActOut ctx) {
  extern __shared__ float smem[];
  const int inner = a.H * 64, c4n = inner >> 2;
  const int m = blockIdx.x, c = threadIdx.x;   // m = b*S + i
  const bool active = c < c4n;
  const int h = c >> 4;
  const int b = m / a.S, i = m - b * a.S;
  float4 q4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) q4 = reinterpret_cast<const float4*>(a.qkv + (int64_t)m * 3 * inner)[c];
  const float* base = a.qkv + (int64_t)b * a.S * 3 * inner;
  const int64_t* mk = a.mask + (int64_t)b * a.S;
  const float* bias = a.bias;
  const int S = a.S;
  const int64_t ld = 3 * inner;
  const float4 o = attn_core(
      q4, h, c, active, a.H, a.S, smem,
      [&](int p) { return reinterpret_cast<const float4*>(base + p * ld + inner); },
      [&](int p) { return reinterpret_cast<const float4*>(base + p * ld + 2 * inner); },
      [&](int p) { return mk[p] != 0; }, [&](int hh, int p) { return bias[hh * (2 * S - 1) + (p - i + S - 1)]; });
  if (active) act_store4(ctx, (int64_t)m * inner + (int64_t)c * 4, o);
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

__global__ void pack_planes_kernel(const float* __restrict__ src, ActOut out, int64_t numel, float scale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < numel) act_store(out, i, src[i] * scale);
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

int launch_rmsnorm(const float* x, const float* w, ActOut out, int64_t rows, int d, float eps, float scale,
                   cudaStream_t s) {
  if (rows == 0) return 0;
  RB_CUDA(launch_pdl(rmsnorm_kernel<false>, dim3(ceil_div(rows, 4)), dim3(128), 0, s, x, w, out, nullptr, rows, d, eps,
                     scale));
  rb::launch_count()++;
  return 0;
}

int launch_rmsnorm_f32(const float* x, const float* w, float* out, int64_t rows, int d, float eps, cudaStream_t s) {
  if (rows == 0) return 0;
  rmsnorm_kernel<true><<<ceil_div(rows, 4), 128, 0, s>>>(x, w, ActOut{nullptr, 0, 0}, out, rows, d, eps, 1.0f);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

static int attn_threads(int H) { return ((H * 16 + 31) / 32) * 32; }

int launch_self_attn_decode(const SelfAttnArgs& a, ActOut ctx, cudaStream_t s) {
  const int threads = attn_threads(a.H);
  RB_REQUIRE(threads <= 1024, "too many heads (%d)", a.H);
  const size_t smem = (size_t)(a.L + a.H * (a.t + 1)) * sizeof(float);
  RB_CUDA(launch_pdl(self_attn_decode_kernel, dim3(a.M), dim3(threads), smem, s, a, ctx));
  rb::launch_count()++;
  return 0;
}

int launch_cross_attn_decode(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s) {
  const int threads = XHG * 32;
  const size_t smem = (size_t)(2 * 32 * XLD + NBC * XHG * 64 + NBC * XHG * a.S) * sizeof(float);
  RB_REQUIRE(smem <= 200 * 1024, "S=%d too large for the cross-attention kernel", a.S);
  static size_t smem_attr = 48 * 1024;
  if (smem > smem_attr) {
    RB_CUDA(cudaFuncSetAttribute(cross_attn_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr = smem;
  }
  RB_REQUIRE(a.M % a.rows_per_query == 0, "row count %d is not a multiple of rows_per_query %d", a.M,
             a.rows_per_query);
  dim3 grid(a.M / a.rows_per_query, ceil_div(a.rows_per_query, NBC), ceil_div(a.H, XHG));
  RB_CUDA(launch_pdl(cross_attn_decode_kernel, grid, dim3(threads), smem, s, a, ctx));
  rb::launch_count()++;
  return 0;
}

int launch_enc_attn(const EncAttnArgs& a, ActOut ctx, cudaStream_t s) {
  const int threads = attn_threads(a.H);
  const size_t smem = (size_t)a.H * a.S * sizeof(float);
  RB_REQUIRE(smem <= 48 * 1024, "H*S=%d too large for the encoder attention kernel", a.H * a.S);
  enc_attn_kernel<<<a.B * a.S, threads, smem, s>>>(a, ctx);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
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

int launch_pack_planes(const float* src, void* dst, int64_t numel, int64_t plane, int mode, float scale,
                       int* overflow, cudaStream_t s) {
  if (numel == 0) return 0;
  pack_planes_kernel<<<ceil_div(numel, 256), 256, 0, s>>>(src, ActOut{dst, plane, mode, overflow}, numel, scale);
  RB_CUDA(cudaGetLastError());
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
