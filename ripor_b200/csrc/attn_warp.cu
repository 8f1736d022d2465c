// Decode-step attention, one WARP per (decoder row, head) with lane = key position.
//
// Same arithmetic as the row-per-CTA kernels in kernels.cu (HF T5 attention as the reference drives it through
// t5_pretrainer/modeling/t5_generative_retriever.py:403-416: no 1/sqrt(dk) scaling, additive relative position
// bias in self-attention only, hard masks, fp32 softmax; the softmax is evaluated chunk by chunk of 32 keys with
// a running maximum). What changes is the shape of the memory traffic. The first versions walked the key
// positions in a serial loop (one dependent L2/HBM round trip per position, ~20 us per CTA); a lane-per-row
// layout is no better (every 128-bit load touches 32 different lines: 512 L1 wavefronts per warp). Here every
// K/V row segment (64 floats = 256 B) is read by 16 lanes as one coalesced 256 B request, all requests of a
// chunk are in flight at once, and
//   self   reduces q.k over the 16 lanes with shuffles (one q per K row), reads (t+1) * 2 * 256 B per (row, head)
//          from the KV cache through the beam ancestry table: HBM-bound
//   cross  stages the chunk in shared memory with cp.async (zero-filled for masked keys) because each K/V row is
//          reused by all beams of the query (the reference expands the encoder states x num_beams,
//          generation.py:231-233; we never do): n_valid * 2 * 256 B per (query, head)
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "kernels.h"

namespace rb {
namespace {

constexpr int kWarps = 4;          // warps per CTA

__device__ __forceinline__ float dot4(float4 a, float4 b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------------------
// self-attention against the KV cache
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWarps * 32, 5) self_attn_warp_kernel(SelfAttnArgs a, ActOut ctx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int wid = blockIdx.x * kWarps + warp;
  const int m = wid / a.H, h = wid - m * a.H;
  pdl_wait();
  if (m >= a.M) return;
  const int inner = a.H * 64, t = a.t, L = a.L;
  const int arow = (a.rpq == 1) ? m * a.nb : m;
  const float* qrow = a.qkv + (int64_t)m * 3 * inner + h * 64;
  {
    const int64_t dst = ((int64_t)t * a.row_cap + m) * inner + h * 64 + lane * 2;
    *reinterpret_cast<float2*>(a.cache_k + dst) = *reinterpret_cast<const float2*>(qrow + inner + lane * 2);
    *reinterpret_cast<float2*>(a.cache_v + dst) = *reinterpret_cast<const float2*>(qrow + 2 * inner + lane * 2);
  }
  const float4 q4 = *reinterpret_cast<const float4*>(qrow + l16 * 4);
  const int64_t hoff = h * 64 + l16 * 4;
  float run_max = -INFINITY, run_sum = 0.f;
  float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c0 = 0; c0 <= t; c0 += 16) {
    const int pl = c0 + l16;
    const int slot_l = pl < t ? pl * (int)a.row_cap + a.anc[(int64_t)arow * L + pl] : -1;
    const float bias_l = pl <= t ? __ldg(a.bias + h * L + (t - pl)) : 0.f;
    float4 k4[8], v4[8];
    float bias[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int slot = __shfl_sync(0xffffffffu, slot_l, 2 * j + half);
      bias[j] = __shfl_sync(0xffffffffu, bias_l, 2 * j + half);
      k4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      v4[j] = k4[j];
      if (c0 + 2 * j + half <= t) {
        const float* kp = slot >= 0 ? a.cache_k + (int64_t)slot * inner + hoff : qrow + inner + l16 * 4;
        const float* vp = slot >= 0 ? a.cache_v + (int64_t)slot * inner + hoff : qrow + 2 * inner + l16 * 4;
        k4[j] = *reinterpret_cast<const float4*>(kp);
        v4[j] = *reinterpret_cast<const float4*>(vp);
      }
    }
    float sc[8];
    float cmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float part = dot4(q4, k4[j], 0.f);
      part += __shfl_xor_sync(0xffffffffu, part, 8);
      part += __shfl_xor_sync(0xffffffffu, part, 4);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      sc[j] = (c0 + 2 * j + half <= t) ? part + bias[j] : -INFINITY;
      cmax = fmaxf(cmax, sc[j]);
    }
    cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, 16));
    const float new_max = fmaxf(run_max, cmax);
    const float rescale = expf(run_max - new_max);
    run_max = new_max;
    o4.x *= rescale; o4.y *= rescale; o4.z *= rescale; o4.w *= rescale;
    float csum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float e = expf(sc[j] - new_max);
      csum += e;
      o4.x = fmaf(e, v4[j].x, o4.x); o4.y = fmaf(e, v4[j].y, o4.y);
      o4.z = fmaf(e, v4[j].z, o4.z); o4.w = fmaf(e, v4[j].w, o4.w);
    }
    run_sum = run_sum * rescale + csum + __shfl_xor_sync(0xffffffffu, csum, 16);
  }
  pdl_trigger();
  o4.x += __shfl_xor_sync(0xffffffffu, o4.x, 16);
  o4.y += __shfl_xor_sync(0xffffffffu, o4.y, 16);
  o4.z += __shfl_xor_sync(0xffffffffu, o4.z, 16);
  o4.w += __shfl_xor_sync(0xffffffffu, o4.w, 16);
  if (half == 0) {
    const float inv = 1.0f / run_sum;
    act_store4(ctx, (int64_t)m * inner + hoff, make_float4(o4.x * inv, o4.y * inv, o4.z * inv, o4.w * inv));
  }
}

// Shared-memory staged variant: ALL K/V rows of a super-chunk (up to 32 positions = 16 KB per warp) are requested
// with cp.async before anything is computed, so a warp's whole working set is in flight at once and costs no
// registers; the dynamic shared memory is sized from t, so early steps (few positions) run many more warps per SM.
// Scores: lane = position (K row from shared memory, 16-byte chunks XOR-swizzled by row so the row-per-lane reads
// are conflict free); context: lane = a pair of output dims.
__device__ __forceinline__ void cp_async16_on(float* dst_smem, const float* src) {
  const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(src) : "memory");
}

__global__ void __launch_bounds__(kWarps * 32) self_attn_smem_kernel(SelfAttnArgs a, ActOut ctx, int pcap) {
  extern __shared__ __align__(16) float ssmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int per_warp = pcap * 128 + 64 + 32;                        // K | V | q | exp(scores)
  float* ks = ssmem + warp * per_warp;
  float* vs = ks + pcap * 64;
  float* qs = vs + pcap * 64;
  float* es = qs + 64;
  const int inner = a.H * 64, t = a.t, L = a.L;
  const int ntask = a.M * a.H;
  pdl_wait();
  // persistent: a few resident CTAs per SM walk the (row, head) tasks (7680 four-warp CTAs cost more in block
  // scheduling than the short early steps take to compute)
  for (int wid = blockIdx.x * kWarps + warp; wid < ntask; wid += gridDim.x * kWarps) {
  const int m = wid / a.H, h = wid - m * a.H;
  const int arow = (a.rpq == 1) ? m * a.nb : m;
  const float* qrow = a.qkv + (int64_t)m * 3 * inner + h * 64;     // q | k | v of this row at position t
  __syncwarp();                                                     // the previous task's readers are done
  if (lane < 16) cp_async16_on(qs + lane * 4, qrow + lane * 4);
  const float* ck = a.cache_k + h * 64 + l16 * 4;
  const float* cv = a.cache_v + h * 64 + l16 * 4;
  float run_max = -INFINITY, run_sum = 0.f;
  float2 o2 = make_float2(0.f, 0.f);
  for (int c0 = 0; c0 <= t; c0 += 32) {
    const int np = min(32, t + 1 - c0);                             // positions of this super-chunk (warp-uniform)
    if (c0 > 0) __syncwarp();
    // ---- stage: lane p owns the cache slot of position c0 + p; 16 lanes x 16 B per row, 2 rows per instruction --
    const int pl = c0 + lane;
    const int slot_l = pl < t ? pl * (int)a.row_cap + a.anc[(int64_t)arow * L + pl] : -1;   // -1: position t (qkv)
    const int nr = (np + 1) >> 1;
#pragma unroll 4
    for (int r = 0; r < nr; ++r) {
      const int row = 2 * r + half;
      const int slot = __shfl_sync(0xffffffffu, slot_l, row);
      if (row < np) {
        const float* kp = slot >= 0 ? ck + (int64_t)slot * inner : qrow + inner + l16 * 4;
        const float* vp = slot >= 0 ? cv + (int64_t)slot * inner : qrow + 2 * inner + l16 * 4;
        cp_async16_on(ks + row * 64 + ((l16 ^ (row & 7)) << 2), kp);
        cp_async16_on(vs + row * 64 + l16 * 4, vp);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    const float bias = lane < np ? __ldg(a.bias + h * L + (t - pl)) : 0.f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- scores: lane = position ---------------------------------------------------------------------------------
    float sc = -INFINITY;
    if (lane < np) {
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        acc = dot4(*reinterpret_cast<const float4*>(qs + j * 4),
                   *reinterpret_cast<const float4*>(ks + lane * 64 + ((j ^ (lane & 7)) << 2)), acc);
      sc = acc + bias;
    }
    const float new_max = fmaxf(run_max, warp_max(sc));
    const float rescale = expf(run_max - new_max);                  // 0 on the first chunk
    const float e = expf(sc - new_max);                             // 0 beyond np
    run_max = new_max;
    run_sum = run_sum * rescale + warp_sum(e);
    o2.x *= rescale;
    o2.y *= rescale;
    es[lane] = e;
    __syncwarp();
    // ---- context: lane = a pair of output dims; 4 positions per iteration (rows beyond np hold e = 0, and their
    //      V rows are zero-filled below so that 0 * garbage never happens) -----------------------------------------
    const int ng = (np + 3) >> 2;
    for (int row = np + half; row < ng * 4; row += 2) *reinterpret_cast<float4*>(vs + row * 64 + l16 * 4) =
        make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
#pragma unroll 2
    for (int g = 0; g < ng; ++g) {
      const float4 e4 = *reinterpret_cast<const float4*>(es + 4 * g);
      const float2 v0 = *reinterpret_cast<const float2*>(vs + (4 * g + 0) * 64 + lane * 2);
      const float2 v1 = *reinterpret_cast<const float2*>(vs + (4 * g + 1) * 64 + lane * 2);
      const float2 v2 = *reinterpret_cast<const float2*>(vs + (4 * g + 2) * 64 + lane * 2);
      const float2 v3 = *reinterpret_cast<const float2*>(vs + (4 * g + 3) * 64 + lane * 2);
      o2.x = fmaf(e4.x, v0.x, o2.x); o2.y = fmaf(e4.x, v0.y, o2.y);
      o2.x = fmaf(e4.y, v1.x, o2.x); o2.y = fmaf(e4.y, v1.y, o2.y);
      o2.x = fmaf(e4.z, v2.x, o2.x); o2.y = fmaf(e4.z, v2.y, o2.y);
      o2.x = fmaf(e4.w, v3.x, o2.x); o2.y = fmaf(e4.w, v3.y, o2.y);
    }
  }
  // this position's K/V go to cache slot (t, m) for the later steps: 64 floats per head, one float2 per lane
  {
    const int64_t dst = ((int64_t)t * a.row_cap + m) * inner + h * 64 + lane * 2;
    *reinterpret_cast<float2*>(a.cache_k + dst) = *reinterpret_cast<const float2*>(qrow + inner + lane * 2);
    *reinterpret_cast<float2*>(a.cache_v + dst) = *reinterpret_cast<const float2*>(qrow + 2 * inner + lane * 2);
  }
  const float inv = 1.0f / run_sum;
  act_store2(ctx, (int64_t)m * inner + h * 64 + lane * 2, make_float2(o2.x * inv, o2.y * inv));
  }
  pdl_trigger();
}

// Forced tail: all T remaining positions of a beam row in one task. The K/V rows of the whole lineage (cached
// prefix + the T new positions) are staged once and serve T queries; the lane that owns key position p keeps its K
// row in registers across the queries.
constexpr int kTailWarpFloats = 32 * 64 * 3 + 32;          // K | V | q rows of up to 32 positions | exp(scores)

__global__ void __launch_bounds__(kWarps * 32, 2) self_attn_tail_kernel(TailAttnArgs a, ActOut ctx) {
  extern __shared__ __align__(16) float tsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  float* ks = tsmem + warp * kTailWarpFloats;
  float* vs = ks + 32 * 64;
  float* qs = vs + 32 * 64;
  float* es = qs + 32 * 64;
  const int inner = a.H * 64, t = a.t, T = a.T, L = a.L, R = a.R;
  const int P = t + T;                                              // positions of the finished lineage (<= 32)
  const int ntask = R * a.H;
  pdl_wait();
  for (int wid = blockIdx.x * kWarps + warp; wid < ntask; wid += gridDim.x * kWarps) {
    const int r = wid / a.H, h = wid - r * a.H;
    __syncwarp();
    // ---- stage K, V of every position and q of the T new ones ------------------------------------------------------
    const int slot_l = lane < t ? lane * (int)a.row_cap + a.anc[(int64_t)r * L + lane] : -1;
    for (int rr = 0; rr < 16; ++rr) {
      const int p = 2 * rr + half;
      const int slot = __shfl_sync(0xffffffffu, slot_l, p);
      if (p < P) {
        const float* kp;
        const float* vp;
        if (p < t) {
          kp = a.cache_k + (int64_t)slot * inner + h * 64 + l16 * 4;
          vp = a.cache_v + (int64_t)slot * inner + h * 64 + l16 * 4;
        } else {
          const float* row = a.qkv + ((int64_t)(p - t) * R + r) * 3 * inner + h * 64 + l16 * 4;
          kp = row + inner;
          vp = row + 2 * inner;
          cp_async16_on(qs + (p - t) * 64 + l16 * 4, row);
        }
        cp_async16_on(ks + p * 64 + ((l16 ^ (p & 7)) << 2), kp);
        cp_async16_on(vs + p * 64 + l16 * 4, vp);
      } else {
        *reinterpret_cast<float4*>(vs + p * 64 + l16 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);   // never 0 * garbage
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    float4 k4[16];                                                  // this lane's key row (position = lane)
#pragma unroll
    for (int j = 0; j < 16; ++j)
      k4[j] = lane < P ? *reinterpret_cast<const float4*>(ks + lane * 64 + ((j ^ (lane & 7)) << 2))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = 0; j < T; ++j) {
      const int pq = t + j;                                         // position of this query
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) acc = dot4(*reinterpret_cast<const float4*>(qs + j * 64 + c * 4), k4[c], acc);
      const float sc = lane <= pq ? acc + __ldg(a.bias + h * L + (pq - lane)) : -INFINITY;
      const float mx = warp_max(sc);
      const float e = expf(sc - mx);                                // 0 beyond the query's own position
      const float sum = warp_sum(e);
      __syncwarp();
      es[lane] = e;
      __syncwarp();
      float2 o2 = make_float2(0.f, 0.f);
      const int ng = (pq + 4) >> 2;
      for (int g = 0; g < ng; ++g) {
        const float4 e4 = *reinterpret_cast<const float4*>(es + 4 * g);
        const float2 v0 = *reinterpret_cast<const float2*>(vs + (4 * g + 0) * 64 + lane * 2);
        const float2 v1 = *reinterpret_cast<const float2*>(vs + (4 * g + 1) * 64 + lane * 2);
        const float2 v2 = *reinterpret_cast<const float2*>(vs + (4 * g + 2) * 64 + lane * 2);
        const float2 v3 = *reinterpret_cast<const float2*>(vs + (4 * g + 3) * 64 + lane * 2);
        o2.x = fmaf(e4.x, v0.x, o2.x); o2.y = fmaf(e4.x, v0.y, o2.y);
        o2.x = fmaf(e4.y, v1.x, o2.x); o2.y = fmaf(e4.y, v1.y, o2.y);
        o2.x = fmaf(e4.z, v2.x, o2.x); o2.y = fmaf(e4.z, v2.y, o2.y);
        o2.x = fmaf(e4.w, v3.x, o2.x); o2.y = fmaf(e4.w, v3.y, o2.y);
      }
      const float inv = 1.0f / sum;
      act_store2(ctx, ((int64_t)j * R + r) * inner + h * 64 + lane * 2, make_float2(o2.x * inv, o2.y * inv));
    }
  }
  pdl_trigger();
}

// ------------------------------------------------------------------------------------------------------------
// cross-attention against the query's encoder K/V
// ------------------------------------------------------------------------------------------------------------
// QS: the beams' q rows are staged in shared memory too; XB: beams (query rows) per warp. Fewer beams per warp =
// fewer registers and less shared memory per warp, i.e. more resident warps to hide the staging latency, at the
// price of staging a query's K/V once per group of XB beams (the repeats hit L2).
template <bool QS, int XB>
constexpr int x_warp_floats() { return 32 * 64 + 32 * 64 + XB * 32 + (QS ? XB * 64 : 0); }   // K | V | exp | q

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool on) {
  const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d32), "l"(src), "r"(on ? 16 : 0) : "memory");
}

template <bool QS, int XB>
__global__ void __launch_bounds__(kWarps * 32, XB <= 5 ? 3 : 2) cross_attn_warp_kernel(CrossAttnArgs a, ActOut ctx) {
  constexpr int kXB = XB;
  extern __shared__ __align__(16) float xsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  float* ks = xsmem + warp * x_warp_floats<QS, XB>();
  float* vs = ks + 32 * 64;                                        // K rows: 16-byte chunks XOR-swizzled by row
  float* es = vs + 32 * 64;
  float* qs = es + kXB * 32;
  const int wid = blockIdx.x * kWarps + warp;
  const int b = wid / a.H, h = wid - b * a.H;
  const int rpq = a.rows_per_query, S = a.S;
  pdl_wait();
  if (b * rpq >= a.M) return;
  const int inner = a.H * 64;
  const int i0 = blockIdx.y * kXB;
  const int nact = min(kXB, rpq - i0);
  const int64_t row0 = (int64_t)blockIdx.z * a.block_rows + (int64_t)b * rpq + i0;
  const int64_t* mk = a.mask + (int64_t)b * S;
  // q rows of this warp's beams -> shared memory (joins the first K commit group); rows beyond nact repeat row 0
  const int64_t q_ld = a.q_ld ? a.q_ld : inner;
  const float* qg = a.q + row0 * q_ld + h * 64;
  if (QS) {
    const float* qb = qg + l16 * 4;
#pragma unroll
    for (int r = 0; r < (kXB + 1) / 2; ++r) {
      const int i = 2 * r + half;
      if (i < kXB) cp_async16(qs + i * 64 + l16 * 4, qb + (int64_t)(i < nact ? i : 0) * q_ld, true);
    }
  }
  // this lane's source pointer for staging: row (half) of the chunk, 16-byte column l16; advances 2 rows per step
  const float* kv_lane = a.kv + ((int64_t)b * S + half) * a.ld + h * 64 + l16 * 4;
  const int64_t step2 = 2 * a.ld;

  float run_max[kXB], run_sum[kXB];
  float2 o2[kXB];
#pragma unroll
  for (int i = 0; i < kXB; ++i) { run_max[i] = -INFINITY; run_sum[i] = 0.f; o2[i] = make_float2(0.f, 0.f); }

  for (int c0 = 0; c0 < S; c0 += 32) {
    const int p = c0 + lane;
    const bool ok = p < S && mk[p < S ? p : 0] != 0;
    const unsigned bits = __ballot_sync(0xffffffffu, ok);
    if (bits == 0) continue;                                       // warp-uniform: a fully masked chunk
    if (c0 > 0) __syncwarp();                                      // previous chunk's readers are done with ks/vs/es
    // ---- stage the chunk: 16 lanes x 16 B per row, two rows per instruction, zero fill for masked keys ---------
    {
      const float* src = kv_lane + (int64_t)c0 * a.ld + a.k_off;
      const unsigned mybits = bits >> half;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const bool on = (mybits >> (2 * r)) & 1u;
        cp_async16(ks + (2 * r + half) * 64 + ((l16 ^ ((2 * r + half) & 7)) << 2), on ? src : a.kv, on);
        src += step2;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      src = kv_lane + (int64_t)c0 * a.ld + a.v_off;
#pragma unroll
      for (int r = 0; r < 16; ++r) {
        const bool on = (mybits >> (2 * r)) & 1u;
        cp_async16(vs + (2 * r + half) * 64 + l16 * 4, on ? src : a.kv, on);
        src += step2;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    // ---- scores: lane = key; the K row comes from shared memory once and meets every beam's q -----------------
    float sc[kXB];
#pragma unroll
    for (int i = 0; i < kXB; ++i) sc[i] = 0.f;
#pragma unroll 4
    for (int j = 0; j < 16; ++j) {
      const float4 k4 = *reinterpret_cast<const float4*>(ks + lane * 64 + ((j ^ (lane & 7)) << 2));
#pragma unroll
      for (int i = 0; i < kXB; ++i) {
        const float4 q4 = QS ? *reinterpret_cast<const float4*>(qs + i * 64 + j * 4)
                             : __ldg(reinterpret_cast<const float4*>(qg + (int64_t)(i < nact ? i : 0) * q_ld) + j);
        sc[i] = dot4(q4, k4, sc[i]);
      }
    }
    if (a.rel_bias != nullptr && ok) {   // encoder: query row i0 + i of the sequence, key p
      const float* rb_h = a.rel_bias + (int64_t)h * (2 * S - 1) + (p + S - 1 - i0);
#pragma unroll
      for (int i = 0; i < kXB; ++i)
        if (i < nact) sc[i] += __ldg(rb_h - i);
    }
#pragma unroll
    for (int i = 0; i < kXB; ++i) {
      const float s_i = ok ? sc[i] : -INFINITY;
      const float new_max = fmaxf(run_max[i], warp_max(s_i));      // finite: the chunk has an unmasked key
      const float rescale = expf(run_max[i] - new_max);
      const float e = expf(s_i - new_max);                         // 0 for masked keys
      run_max[i] = new_max;
      run_sum[i] = run_sum[i] * rescale + warp_sum(e);
      o2[i].x *= rescale;
      o2[i].y *= rescale;
      es[i * 32 + lane] = e;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // ---- context: lane = a pair of output dims; 4 keys per iteration from shared memory ------------------------
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (((bits >> (4 * g)) & 0xFu) == 0) continue;               // warp-uniform
      const float2 v0 = *reinterpret_cast<const float2*>(vs + (4 * g + 0) * 64 + lane * 2);
      const float2 v1 = *reinterpret_cast<const float2*>(vs + (4 * g + 1) * 64 + lane * 2);
      const float2 v2 = *reinterpret_cast<const float2*>(vs + (4 * g + 2) * 64 + lane * 2);
      const float2 v3 = *reinterpret_cast<const float2*>(vs + (4 * g + 3) * 64 + lane * 2);
#pragma unroll
      for (int i = 0; i < kXB; ++i) {
        const float4 e4 = *reinterpret_cast<const float4*>(es + i * 32 + 4 * g);
        o2[i].x = fmaf(e4.x, v0.x, o2[i].x); o2[i].y = fmaf(e4.x, v0.y, o2[i].y);
        o2[i].x = fmaf(e4.y, v1.x, o2[i].x); o2[i].y = fmaf(e4.y, v1.y, o2[i].y);
        o2[i].x = fmaf(e4.z, v2.x, o2[i].x); o2[i].y = fmaf(e4.z, v2.y, o2[i].y);
        o2[i].x = fmaf(e4.w, v3.x, o2[i].x); o2[i].y = fmaf(e4.w, v3.y, o2[i].y);
      }
    }
  }
  pdl_trigger();
  asm volatile("cp.async.wait_group 0;" ::: "memory");             // (a fully masked query never waited for its q rows)
#pragma unroll
  for (int i = 0; i < kXB; ++i)
    if (i < nact) {
      const float inv = run_sum[i] > 0.f ? 1.0f / run_sum[i] : 0.f;
      act_store2(ctx, (row0 + i) * inner + h * 64 + lane * 2, make_float2(o2[i].x * inv, o2[i].y * inv));
    }
}

}  // namespace

bool launch_self_attn_warp(const SelfAttnArgs& a, ActOut ctx, cudaStream_t s, int* status) {
  const dim3 grid(ceil_div((int64_t)a.M * a.H, kWarps)), block(kWarps * 32);
  // Staged kernel while few positions are cached (its shared memory grows with t and the early steps are bound
  // by per-task latency, so occupancy wins); register kernel for the long, HBM-bound steps. Measured crossover on
  // B200 at 24 positions. RB200_SELF=smem|reg forces one of them.
  static const int force = []() {
    const char* e = getenv("RB200_SELF");
    return !e ? 0 : (strcmp(e, "smem") == 0 ? 1 : (strcmp(e, "reg") == 0 ? 2 : 0));
  }();
  const bool staged = force == 1 || (force == 0 && a.t + 1 <= 24);
  cudaError_t err;
  if (staged) {
    const int pcap = ((a.t + 1 < 32 ? a.t + 1 : 32) + 3) & ~3;      // staged positions per warp, multiple of 4
    const size_t smem = (size_t)kWarps * (pcap * 128 + 64 + 32) * sizeof(float);
    static bool attr_set = false;
    static int sms = 0;
    if (!attr_set) {
      err = cudaFuncSetAttribute(self_attn_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)kWarps * (32 * 128 + 96) * sizeof(float)));
      int dev = 0;
      if (err == cudaSuccess) err = cudaGetDevice(&dev);
      if (err == cudaSuccess) err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (err != cudaSuccess) {
        *status = fail(RB200_ERR_CUDA, "self_attn_smem_kernel attribute: %s", cudaGetErrorString(err));
        return true;
      }
      attr_set = true;
    }
    int per_sm = (int)((220 * 1024) / (smem + 1024));
    per_sm = per_sm > 7 ? 7 : (per_sm < 1 ? 1 : per_sm);           // 72 registers x 128 threads: 7 CTAs per SM
    const int want = ceil_div((int64_t)a.M * a.H, kWarps);
    const dim3 pgrid(want < sms * per_sm ? want : sms * per_sm);
    err = launch_pdl(self_attn_smem_kernel, pgrid, block, smem, s, a, ctx, pcap);
  } else {
    err = launch_pdl(self_attn_warp_kernel, grid, block, 0, s, a, ctx);
  }
  *status = err == cudaSuccess ? 0 : fail(RB200_ERR_CUDA, "self_attn_warp_kernel launch: %s", cudaGetErrorString(err));
  launch_count()++;
  return true;
}

template <bool QS, int XB>
static cudaError_t launch_cross_cfg(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s) {
  const int B = a.M / a.rows_per_query;
  const dim3 grid(ceil_div((int64_t)B * a.H, kWarps), ceil_div(a.rows_per_query, XB), a.nblocks), block(kWarps * 32);
  constexpr size_t smem = (size_t)kWarps * x_warp_floats<QS, XB>() * sizeof(float);
  auto kern = cross_attn_warp_kernel<QS, XB>;
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_pdl(kern, grid, block, smem, s, a, ctx);
}

int launch_self_attn_tail(const TailAttnArgs& a, ActOut ctx, cudaStream_t s) {
  RB_REQUIRE(a.t >= 1 && a.T >= 1 && a.t + a.T <= 32, "forced tail needs 1 <= t and t + T <= 32 (t=%d, T=%d)", a.t, a.T);
  constexpr size_t smem = (size_t)kWarps * kTailWarpFloats * sizeof(float);
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    RB_CUDA(cudaFuncSetAttribute(self_attn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  const int want = ceil_div((int64_t)a.R * a.H, kWarps);
  const dim3 grid(want < 2 * sms ? want : 2 * sms), block(kWarps * 32);
  RB_CUDA(launch_pdl(self_attn_tail_kernel, grid, block, smem, s, a, ctx));
  launch_count()++;
  return 0;
}

bool launch_cross_attn_warp(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s, int* status) {
  static const bool qs = []() {
    const char* e = getenv("RB200_XATTN_Q");     // default: q rows staged in shared memory (24 us vs 30 us per launch)
    return !(e && strcmp(e, "ldg") == 0);
  }();
  static const int xb_env = []() {
    const char* e = getenv("RB200_XATTN_B");
    return e ? atoi(e) : 0;
  }();
  // 5 beams per warp up to 20 rows per query (the bench shape: 10), 10 beyond (beam 100: fewer K/V restagings)
  const int xb = xb_env ? xb_env : (a.rows_per_query <= 20 ? 5 : 10);
  cudaError_t err;
  if (xb == 5) err = qs ? launch_cross_cfg<true, 5>(a, ctx, s) : launch_cross_cfg<false, 5>(a, ctx, s);
  else err = qs ? launch_cross_cfg<true, 10>(a, ctx, s) : launch_cross_cfg<false, 10>(a, ctx, s);
  *status = err == cudaSuccess ? 0 : fail(RB200_ERR_CUDA, "cross_attn_warp_kernel launch: %s", cudaGetErrorString(err));
  launch_count()++;
  return true;
}

}  // namespace rb
