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
// compact row m of the step loop -> original row id (query id when one row per query runs, i.e. step 0)
__device__ __forceinline__ int orig_row(const SelfAttnArgs& a, int m) {
  if (a.qlist == nullptr) return m;
  return a.rpq == 1 ? a.qlist[m] : a.qlist[m / a.nb] * a.nb + m % a.nb;
}

__global__ void __launch_bounds__(kWarps * 32, 5) self_attn_warp_kernel(SelfAttnArgs a, ActOut ctx) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  const int wid = blockIdx.x * kWarps + warp;
  const int m = wid / a.H, h = wid - m * a.H;
  pdl_wait();
  if (m >= a.M) return;
  const int inner = a.H * 64, t = a.t, L = a.L;
  const int orow = orig_row(a, m);                                  // cache slots / ancestry: original row ids
  const int arow = (a.rpq == 1) ? orow * a.nb : orow;
  const float* qrow = a.qkv + (int64_t)m * 3 * inner + h * 64;
  {
    const int64_t dst = ((int64_t)t * a.row_cap + orow) * inner + h * 64 + lane * 2;
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

__global__ void __launch_bounds__(kWarps * 32, 7) self_attn_smem_kernel(SelfAttnArgs a, ActOut ctx, int pcap) {
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
  // The original row id and the lineage's cache slots of a task come out of two dependent loads (compaction list ->
  // ancestry table) that the K/V requests depend on in turn: they are fetched one task ahead.
  const int stride = gridDim.x * kWarps;
  auto task_rows = [&](int wid, int& orow, int& slot0) {
    const int m = wid / a.H;
    orow = orig_row(a, m);
    const int arow = (a.rpq == 1) ? orow * a.nb : orow;
    slot0 = lane < t ? lane * (int)a.row_cap + a.anc[(int64_t)arow * L + lane] : -1;     // positions 0..31 (chunk 0)
  };
  int orow_n = 0, slot0_n = -1;
  if ((int)(blockIdx.x * kWarps + warp) < ntask) task_rows(blockIdx.x * kWarps + warp, orow_n, slot0_n);
  for (int wid = blockIdx.x * kWarps + warp; wid < ntask; wid += stride) {
  const int m = wid / a.H, h = wid - m * a.H;
  const int orow = orow_n, slot0 = slot0_n;
  const int arow = (a.rpq == 1) ? orow * a.nb : orow;
  if (wid + stride < ntask) task_rows(wid + stride, orow_n, slot0_n);
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
    const int slot_l = c0 == 0 ? slot0
                               : (pl < t ? pl * (int)a.row_cap + a.anc[(int64_t)arow * L + pl] : -1);   // -1: position t (qkv)
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
    const int64_t dst = ((int64_t)t * a.row_cap + orow) * inner + h * 64 + lane * 2;
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
constexpr int kTailJB = 4;                                  // queries in flight per warp (independent chains)
constexpr int kTailWarpFloats = 32 * 64 * 3 + kTailJB * 32;   // K | V | q rows of up to 32 positions | exp(scores)

__global__ void __launch_bounds__(kWarps * 32, 2) self_attn_tail_kernel(TailAttnArgs a, ActOut ctx) {
  extern __shared__ __align__(16) float tsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  float* ks = tsmem + warp * kTailWarpFloats;
  float* vs = ks + 32 * 64;
  float* qs = vs + 32 * 64;
  float* es = qs + 32 * 64;
  const int inner = a.H * 64, L = a.L;
  const int P = a.lay.P;                                            // positions of the finished lineage (<= 32)
  const int ntask = a.R * a.H;
  pdl_wait();
  for (int wid = blockIdx.x * kWarps + warp; wid < ntask; wid += gridDim.x * kWarps) {
    const int rp = wid / a.H, h = wid - rp * a.H;                   // frozen row (freeze order), head
    const int bq = a.fz_list[rp / a.nb];
    const int r = bq * a.nb + rp % a.nb;                            // original row id
    const int t = a.qstart ? a.qstart[bq] : 0, T = P - t;           // this row's pass: positions t..P-1
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
          const float* row = a.qkv + ((int64_t)a.lay.off[p] + rp) * 3 * inner + h * 64 + l16 * 4;
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
    // kTailJB queries at a time: their dot-product, shuffle and accumulate chains are independent, which is what
    // keeps the few resident warps (shared memory: 2 CTAs per SM) issuing
    for (int j0 = 0; j0 < T; j0 += kTailJB) {
      float acc[kTailJB];
#pragma unroll
      for (int u = 0; u < kTailJB; ++u) acc[u] = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {                                // fully unrolled: k4[] must stay in registers
#pragma unroll
        for (int u = 0; u < kTailJB; ++u)
          acc[u] = dot4(*reinterpret_cast<const float4*>(qs + min(j0 + u, T - 1) * 64 + c * 4), k4[c], acc[u]);
      }
      float inv[kTailJB];
      __syncwarp();                                                 // the previous group's readers of es are done
#pragma unroll
      for (int u = 0; u < kTailJB; ++u) {
        const int pq = t + min(j0 + u, T - 1);                      // position of this query
        const float sc = lane <= pq ? acc[u] + __ldg(a.bias + h * L + (pq - lane)) : -INFINITY;
        const float mx = warp_max(sc);
        const float e = expf(sc - mx);                              // 0 beyond the query's own position
        inv[u] = 1.0f / warp_sum(e);
        es[u * 32 + lane] = e;
      }
      __syncwarp();
      float2 o2[kTailJB];
#pragma unroll
      for (int u = 0; u < kTailJB; ++u) o2[u] = make_float2(0.f, 0.f);
      const int ng = (t + min(j0 + kTailJB - 1, T - 1) + 4) >> 2;   // groups of 4 positions up to the last query's own
      for (int g = 0; g < ng; ++g) {
        const float2 v0 = *reinterpret_cast<const float2*>(vs + (4 * g + 0) * 64 + lane * 2);
        const float2 v1 = *reinterpret_cast<const float2*>(vs + (4 * g + 1) * 64 + lane * 2);
        const float2 v2 = *reinterpret_cast<const float2*>(vs + (4 * g + 2) * 64 + lane * 2);
        const float2 v3 = *reinterpret_cast<const float2*>(vs + (4 * g + 3) * 64 + lane * 2);
#pragma unroll
        for (int u = 0; u < kTailJB; ++u) {
          const float4 e4 = *reinterpret_cast<const float4*>(es + u * 32 + 4 * g);
          o2[u].x = fmaf(e4.x, v0.x, o2[u].x); o2[u].y = fmaf(e4.x, v0.y, o2[u].y);
          o2[u].x = fmaf(e4.y, v1.x, o2[u].x); o2[u].y = fmaf(e4.y, v1.y, o2[u].y);
          o2[u].x = fmaf(e4.z, v2.x, o2[u].x); o2[u].y = fmaf(e4.z, v2.y, o2[u].y);
          o2[u].x = fmaf(e4.w, v3.x, o2[u].x); o2[u].y = fmaf(e4.w, v3.y, o2[u].y);
        }
      }
#pragma unroll
      for (int u = 0; u < kTailJB; ++u)
        if (j0 + u < T)
          act_store2(ctx, ((int64_t)a.lay.off[t + j0 + u] + rp) * inner + h * 64 + lane * 2,
                     make_float2(o2[u].x * inv[u], o2[u].y * inv[u]));
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
constexpr int x_warp_floats() { return 32 * 64 + 32 * 64 + XB * 32 + (QS ? 2 * XB * 64 : 0); }   // K | V | exp | 2 x q
// ~20 KB of shared memory per warp: ONE CTA per SM with as many warps as fit (9-11) uses the 227 KB better than
// 2 CTAs x 4 warps
template <bool QS, int XB>
constexpr int x_warps() {
  constexpr int n = (226 * 1024) / (x_warp_floats<QS, XB>() * 4);
  return n > 12 ? 12 : n;
}

__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool on) {
  const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d32), "l"(src), "r"(on ? 16 : 0) : "memory");
}

template <bool QS, int XB>
__global__ void __launch_bounds__(x_warps<QS, XB>() * 32, 1) cross_attn_warp_kernel(CrossAttnArgs a, ActOut ctx) {
  constexpr int kXB = XB;
  constexpr int kXWarps = x_warps<QS, XB>();
  extern __shared__ __align__(16) float xsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  float* ks = xsmem + warp * x_warp_floats<QS, XB>();
  float* vs = ks + 32 * 64;                                        // K rows: 16-byte chunks XOR-swizzled by row
  float* es = vs + 32 * 64;
  float* qs = es + kXB * 32;                                       // two buffers of XB rows (QS only)
  const int wid = blockIdx.x * kXWarps + warp;
  const int b = wid / a.H, h = wid - b * a.H;                      // query slot (compact / freeze order), head
  const int rpq = a.rows_per_query, S = a.S;
  pdl_wait();
  if (b * rpq >= a.M) return;
  const int bq = a.qmap ? a.qmap[b] : b;                           // the query whose encoder K/V and mask to use
  const int t0 = (a.ragged && a.qstart) ? a.qstart[bq] : 0;
  const int nblocks = a.ragged ? a.lay.P - t0 : a.nblocks;
  const int inner = a.H * 64;
  const int i0 = blockIdx.y * kXB;
  const int nact = min(kXB, rpq - i0);
  const int64_t* mk = a.mask + (int64_t)bq * S;
  const int64_t q_ld = a.q_ld ? a.q_ld : inner;
  // this lane's source pointer for staging: row (half) of the chunk, 16-byte column l16; advances 2 rows per step
  const float* kv_lane = a.kv + ((int64_t)bq * S + half) * a.ld + h * 64 + l16 * 4;
  const int64_t step2 = 2 * a.ld;
  const bool resident = S <= 32;       // one chunk: the staged K/V serve every position block of the forced tail
  auto row_of = [&](int z) {
    return (a.ragged ? (int64_t)a.lay.off[t0 + z] : (int64_t)z * a.block_rows) + (int64_t)b * rpq + i0;
  };
  auto stage_q = [&](int z, float* dst) {                          // q rows of block z; rows beyond nact repeat row 0
    const float* qb = a.q + row_of(z) * q_ld + h * 64 + l16 * 4;
#pragma unroll
    for (int r = 0; r < (kXB + 1) / 2; ++r) {
      const int i = 2 * r + half;
      if (i < kXB) cp_async16(dst + i * 64 + l16 * 4, qb + (int64_t)(i < nact ? i : 0) * q_ld, true);
    }
  };
  const int z0 = blockIdx.z, zs = gridDim.z;
  if (QS && z0 < nblocks) stage_q(z0, qs);
  asm volatile("cp.async.commit_group;" ::: "memory");
  int qbuf = 0;
  for (int z = z0; z < nblocks; z += zs, qbuf ^= 1) {
    const int64_t row0 = row_of(z);
    const float* qcur = qs + qbuf * (kXB * 64);
    const float* qg = a.q + row0 * q_ld + h * 64;
    __syncwarp();                                                  // the previous block's readers are done
    if (QS && z + zs < nblocks) stage_q(z + zs, qs + (qbuf ^ 1) * (kXB * 64));     // prefetch the next block's q
    asm volatile("cp.async.commit_group;" ::: "memory");           // (possibly empty: uniform group accounting)
    float run_max[kXB], run_sum[kXB];
    float2 o2[kXB];
#pragma unroll
    for (int i = 0; i < kXB; ++i) { run_max[i] = -INFINITY; run_sum[i] = 0.f; o2[i] = make_float2(0.f, 0.f); }
    for (int c0 = 0; c0 < S; c0 += 32) {
      const int p = c0 + lane;
      const bool ok = p < S && mk[p < S ? p : 0] != 0;
      const unsigned bits = __ballot_sync(0xffffffffu, ok);
      if (bits == 0) continue;                                     // warp-uniform: a fully masked chunk
      const bool stage_now = !(resident && z > z0);
      if (stage_now) {
        if (c0 > 0) __syncwarp();                                  // previous chunk's readers are done with ks/vs/es
        // ---- stage the chunk: 16 lanes x 16 B per row, two rows per instruction, zero fill for masked keys -------
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
      asm volatile("cp.async.wait_group 1;" ::: "memory");         // everything but the newest group (V, or next q)
      __syncwarp();
      // ---- scores: lane = key; the K row comes from shared memory once and meets every beam's q ---------------
      float sc[kXB];
#pragma unroll
      for (int i = 0; i < kXB; ++i) sc[i] = 0.f;
#pragma unroll 4
      for (int j = 0; j < 16; ++j) {
        const float4 k4 = *reinterpret_cast<const float4*>(ks + lane * 64 + ((j ^ (lane & 7)) << 2));
#pragma unroll
        for (int i = 0; i < kXB; ++i) {
          const float4 q4 = QS ? *reinterpret_cast<const float4*>(qcur + i * 64 + j * 4)
                               : __ldg(reinterpret_cast<const float4*>(qg + (int64_t)(i < nact ? i : 0) * q_ld) + j);
          sc[i] = dot4(q4, k4, sc[i]);
        }
      }
      if (a.rel_bias != nullptr && ok) {   // encoder: query row i0 + i of the sequence, key p
        const float* rb_h = a.rel_bias + (int64_t)h * (2 * a.rel_S - 1) + (p + a.rel_S - 1 - i0);
#pragma unroll
        for (int i = 0; i < kXB; ++i)
          if (i < nact) sc[i] += __ldg(rb_h - i);
      }
      if (!stage_now) __syncwarp();                                // es of the previous block has been consumed
#pragma unroll
      for (int i = 0; i < kXB; ++i) {
        const float s_i = ok ? sc[i] : -INFINITY;
        const float new_max = fmaxf(run_max[i], warp_max(s_i));    // finite: the chunk has an unmasked key
        const float rescale = expf(run_max[i] - new_max);
        const float e = expf(s_i - new_max);                       // 0 for masked keys
        run_max[i] = new_max;
        run_sum[i] = run_sum[i] * rescale + warp_sum(e);
        o2[i].x *= rescale;
        o2[i].y *= rescale;
        es[i * 32 + lane] = e;
      }
      if (stage_now) asm volatile("cp.async.wait_group 0;" ::: "memory");   // V (and the prefetched q) landed
      __syncwarp();
      // ---- context: lane = a pair of output dims; 4 keys per iteration from shared memory ----------------------
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        if (((bits >> (4 * g)) & 0xFu) == 0) continue;             // warp-uniform
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
#pragma unroll
    for (int i = 0; i < kXB; ++i)
      if (i < nact) {
        const float inv = run_sum[i] > 0.f ? 1.0f / run_sum[i] : 0.f;
        act_store2(ctx, (row0 + i) * inner + h * 64 + lane * 2, make_float2(o2[i].x * inv, o2[i].y * inv));
      }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");             // (a fully masked query never waited for its q rows)
  pdl_trigger();
}

// ------------------------------------------------------------------------------------------------------------
// forced tail, cross-attention on the (legacy) warp-level tensor cores
// ------------------------------------------------------------------------------------------------------------
// In the forced tail every (query, head) pair serves T * nb rows (280 at the bench shape) against the same <= 32
// keys, which makes it a small GEMM pair: S = Q K^T (rows x 32 keys over 64 dims), O = softmax(S) V. The FFMA kernel
// above is issue-bound there (1440 warp instructions per 5 rows). This kernel runs 16-row tiles through
// mma.sync.m16n8k8 tf32 with the same error-compensated 3-product split as the GEMMs (hi*hi + hi*lo + lo*hi, fp32
// accumulate), so the scores keep fp32-grade accuracy. K and V are split once per CTA into shared memory.
constexpr int kMmaLd = 68;   // row stride (floats) of the staged K/V planes: conflict-free fragment reads

__device__ __forceinline__ void mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
        "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

__global__ void __launch_bounds__(128, 3) cross_attn_mma_kernel(CrossAttnArgs a, ActOut ctx) {
  __shared__ __align__(16) float k_hi[32 * kMmaLd], k_lo[32 * kMmaLd], v_hi[32 * kMmaLd], v_lo[32 * kMmaLd];
  __shared__ unsigned valid_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;        // query slot, head
  const int S = a.S, rpq = a.rows_per_query, inner = a.H * 64;
  pdl_wait();
  const int bq = a.qmap ? a.qmap[b] : b;
  const int t0 = (a.ragged && a.qstart) ? a.qstart[bq] : 0;
  // ---- K, V of this (query, head): split into tf32 planes once; masked / absent keys are zero rows --------------
  const int64_t* mk = a.mask + (int64_t)bq * S;
  if (warp == 0) {
    const bool ok = lane < S && mk[lane < S ? lane : 0] != 0;
    const unsigned bits = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) valid_s = bits;
  }
  __syncthreads();
  const unsigned valid = valid_s;
  const float* kvb = a.kv + (int64_t)bq * S * a.ld + h * 64;
  for (int e = threadIdx.x; e < 32 * 16; e += blockDim.x) {
    const int key = e >> 4, c4 = (e & 15) * 4;
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
    if ((valid >> key) & 1u) {
      kk = __ldg(reinterpret_cast<const float4*>(kvb + (int64_t)key * a.ld + a.k_off + c4));
      vv = __ldg(reinterpret_cast<const float4*>(kvb + (int64_t)key * a.ld + a.v_off + c4));
    }
    const float kf[4] = {kk.x, kk.y, kk.z, kk.w}, vf[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float kh = round_tf32(kf[u]), vh = round_tf32(vf[u]);
      k_hi[key * kMmaLd + c4 + u] = kh;
      k_lo[key * kMmaLd + c4 + u] = round_tf32(kf[u] - kh);
      v_hi[key * kMmaLd + c4 + u] = vh;
      v_lo[key * kMmaLd + c4 + u] = round_tf32(vf[u] - vh);
    }
  }
  __syncthreads();
  // ---- 16-row tiles of this query's rows: tile row q <-> (position block q / rpq, beam q % rpq) ------------------
  const int nrows = (a.ragged ? a.lay.P - t0 : a.nblocks) * rpq;
  const int64_t q_ld = a.q_ld ? a.q_ld : inner;
  auto global_row = [&](int q) {
    return (a.ragged ? (int64_t)a.lay.off[t0 + q / rpq] : (int64_t)(q / rpq) * a.block_rows) + (int64_t)b * rpq + (q % rpq);
  };
  // gridDim.y CTAs share one (query, head): each restages the <= 32 keys and takes every gridDim.y-th group of 4 tiles
  for (int tile = blockIdx.y * 4 + warp; tile * 16 < nrows; tile += 4 * gridDim.y) {
    const int q0 = tile * 16 + g, q1 = q0 + 8;                      // this lane's two rows of the tile
    const float* qr0 = a.q + global_row(min(q0, nrows - 1)) * q_ld + h * 64;
    const float* qr1 = a.q + global_row(min(q1, nrows - 1)) * q_ld + h * 64;
    float qa[8][4];                                                 // A fragments of the 8 k-steps (raw fp32)
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      qa[kk][0] = __ldg(qr0 + kk * 8 + t);
      qa[kk][1] = __ldg(qr1 + kk * 8 + t);
      qa[kk][2] = __ldg(qr0 + kk * 8 + t + 4);
      qa[kk][3] = __ldg(qr1 + kk * 8 + t + 4);
    }
    // S = Q K^T: 4 key tiles of 8
    float sacc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int u = 0; u < 4; ++u) sacc[nt][u] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      float ah[4], al[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ah[u] = round_tf32(qa[kk][u]);
        al[u] = round_tf32(qa[kk][u] - ah[u]);
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int o0 = (nt * 8 + g) * kMmaLd + kk * 8 + t;          // B(k = t, n = g) = K[key nt*8+g][dim kk*8+t]
        const float bh0 = k_hi[o0], bh1 = k_hi[o0 + 4], bl0 = k_lo[o0], bl1 = k_lo[o0 + 4];
        mma_tf32(sacc[nt], al, bh0, bh1);
        mma_tf32(sacc[nt], ah, bl0, bl1);
        mma_tf32(sacc[nt], ah, bh0, bh1);
      }
    }
    // softmax over the keys of rows q0 (c0, c1) and q1 (c2, c3); key of c(2u'+w) in tile nt = nt*8 + 2t + w
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const bool on = (valid >> (nt * 8 + 2 * t + w)) & 1u;
        sacc[nt][w] = on ? sacc[nt][w] : -INFINITY;
        sacc[nt][2 + w] = on ? sacc[nt][2 + w] : -INFINITY;
        mx0 = fmaxf(mx0, sacc[nt][w]);
        mx1 = fmaxf(mx1, sacc[nt][2 + w]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        sacc[nt][w] = mx0 == -INFINITY ? 0.f : expf(sacc[nt][w] - mx0);
        sacc[nt][2 + w] = mx1 == -INFINITY ? 0.f : expf(sacc[nt][2 + w] - mx1);
        sum0 += sacc[nt][w];
        sum1 += sacc[nt][2 + w];
      }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // O = P V. The C fragment of key tile ks IS an A fragment if A's column c stands for key 2c (c < 4) / 2(c-4)+1:
    // a0 = c0, a1 = c2, a2 = c1, a3 = c3; the V rows of the B fragment follow the same key order.
    float oacc[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int u = 0; u < 4; ++u) oacc[nd][u] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const float pf[4] = {sacc[ks][0], sacc[ks][2], sacc[ks][1], sacc[ks][3]};
      float ph[4], pl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        ph[u] = round_tf32(pf[u]);
        pl[u] = round_tf32(pf[u] - ph[u]);
      }
#pragma unroll
      for (int nd = 0; nd < 8; ++nd) {
        const int o0 = (ks * 8 + 2 * t) * kMmaLd + nd * 8 + g;       // B(k = t, n = g) = V[key ks*8+2t][dim nd*8+g]
        const float bh0 = v_hi[o0], bh1 = v_hi[o0 + kMmaLd], bl0 = v_lo[o0], bl1 = v_lo[o0 + kMmaLd];
        mma_tf32(oacc[nd], pl, bh0, bh1);
        mma_tf32(oacc[nd], ph, bl0, bl1);
        mma_tf32(oacc[nd], ph, bh0, bh1);
      }
    }
    const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f, inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;
    if (q0 < nrows) {
      const int64_t base = global_row(q0) * inner + h * 64 + 2 * t;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
        act_store2(ctx, base + nd * 8, make_float2(oacc[nd][0] * inv0, oacc[nd][1] * inv0));
    }
    if (q1 < nrows) {
      const int64_t base = global_row(q1) * inner + h * 64 + 2 * t;
#pragma unroll
      for (int nd = 0; nd < 8; ++nd)
        act_store2(ctx, base + nd * 8, make_float2(oacc[nd][2] * inv1, oacc[nd][3] * inv1));
    }
  }
  pdl_trigger();
}

// fp16x3 flavour of the tensor-core cross-attention (precision mode fp16x3 only): the operands are split into two
// fp16 planes like the GEMM operands of that mode (11-bit mantissas, range-checked), which lets m16n8k16 do twice the
// contraction per instruction and halves the fragment traffic. K and V are staged row-major as fp16 planes; the V
// fragments of P V come out of ldmatrix.trans.
constexpr int kKLd = 72;   // halfs per staged K row (64 dims + pad): conflict-free B fragments

// (not volatile: a pure function of its operands, so independent accumulator chains may be interleaved)
__device__ __forceinline__ void mma_f16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// four transposed 8x8 b16 matrices: with V kept row-major [key][dim], lane l passes the address of row
// (key0 + (l & 7) + 8 * ((l >> 3) & 1)) at dim0 + 8 * (l >> 4) and receives b0, b1 of dim tile 0 and b0, b1 of dim tile 1
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const __half* row_ptr) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(row_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
// (x, y) -> fp16 pair hi and the fp16 pair of the remainders; flags values outside the fp16 range
__device__ __forceinline__ void split_h2(float x, float y, uint32_t& hi, uint32_t& lo, bool& bad) {
  bad |= !(fabsf(x) <= kFp16Limit && fabsf(y) <= kFp16Limit);
  const __half2 h = __floats2half2_rn(x, y);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// 4 x 4 transpose across the lanes of a quad (t4 = lane & 3): lane d ends with r[s] = (lane s's r[d])
__device__ __forceinline__ void quad_transpose4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, int t4) {
  const bool odd = t4 & 1, up = t4 & 2;
  uint32_t x = __shfl_xor_sync(0xffffffffu, odd ? r0 : r1, 1);
  if (odd) r0 = x; else r1 = x;
  x = __shfl_xor_sync(0xffffffffu, odd ? r2 : r3, 1);
  if (odd) r2 = x; else r3 = x;
  x = __shfl_xor_sync(0xffffffffu, up ? r0 : r2, 2);
  if (up) r0 = x; else r2 = x;
  x = __shfl_xor_sync(0xffffffffu, up ? r1 : r3, 2);
  if (up) r1 = x; else r3 = x;
}

// One output row of a 16-row m16n8 accumulator tile (64 dims: lane (g, t4) holds columns nd * 8 + 2 * t4, + 1 of the
// eight dim tiles nd) scaled by inv and written as fp16 hi / lo planes. The pairs are split in registers, transposed
// across the quad so that lane t4 owns the whole 16-byte chunks nd = t4 and t4 + 4, and stored with four 16-byte
// stores per row (the generic act_store2 path costs 32 four-byte stores plus a mode switch and a range-check branch
// per pair: ~30 % of the instructions of the tail attention kernels). Must be called by all 32 lanes; `live` guards
// only the stores. row_hi = hi plane of the row at the head's first dim, plane = distance to the lo plane.
__device__ __forceinline__ void store_row_planes_f16(__half* row_hi, int64_t plane, const float (&o)[8][4], int w0,
                                                     float inv, bool live, int t4, bool& bad) {
  uint32_t hi[8], lo[8];
#pragma unroll
  // (no range check: a row of softmax(S) V is a convex combination of V rows that passed theirs when they were split)
  bool unchecked = false;
#pragma unroll
  for (int nd = 0; nd < 8; ++nd) split_h2(o[nd][w0] * inv, o[nd][w0 + 1] * inv, hi[nd], lo[nd], unchecked);
  quad_transpose4(hi[0], hi[1], hi[2], hi[3], t4);
  quad_transpose4(hi[4], hi[5], hi[6], hi[7], t4);
  quad_transpose4(lo[0], lo[1], lo[2], lo[3], t4);
  quad_transpose4(lo[4], lo[5], lo[6], lo[7], t4);
  if (live) {
    *reinterpret_cast<uint4*>(row_hi + t4 * 8) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(row_hi + 32 + t4 * 8) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
    *reinterpret_cast<uint4*>(row_hi + plane + t4 * 8) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    *reinterpret_cast<uint4*>(row_hi + plane + 32 + t4 * 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
  }
}

template <int MINB>   // resident CTAs per SM the register allocation aims at (3: 170 registers, 4: 128)
__global__ void __launch_bounds__(128, MINB) cross_attn_mma16_kernel(CrossAttnArgs a, ActOut ctx) {
  __shared__ __align__(16) __half k_hi[32 * kKLd], k_lo[32 * kKLd], v_hi[32 * kKLd], v_lo[32 * kKLd];
  __shared__ unsigned valid_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.x / a.H, h = blockIdx.x - b * a.H;        // query slot, head
  const int S = a.S, rpq = a.rows_per_query, inner = a.H * 64;
  bool bad = false;
  pdl_wait();
  const int bq = a.qmap ? a.qmap[b] : b;
  const int t0 = (a.ragged && a.qstart) ? a.qstart[bq] : 0;
  const int64_t* mk = a.mask + (int64_t)bq * S;
  if (warp == 0) {
    const bool ok = lane < S && mk[lane < S ? lane : 0] != 0;
    const unsigned bits = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) valid_s = bits;
  }
  __syncthreads();
  const unsigned valid = valid_s;
  const float* kvb = a.kv + (int64_t)bq * S * a.ld + h * 64;
  for (int e = threadIdx.x; e < 32 * 16; e += blockDim.x) {
    const int key = e >> 4, c4 = (e & 15) * 4;
    float4 kk = make_float4(0.f, 0.f, 0.f, 0.f), vv = kk;
    if ((valid >> key) & 1u) {
      kk = __ldg(reinterpret_cast<const float4*>(kvb + (int64_t)key * a.ld + a.k_off + c4));
      vv = __ldg(reinterpret_cast<const float4*>(kvb + (int64_t)key * a.ld + a.v_off + c4));
    }
    uint32_t h0, l0, h1, l1;
    split_h2(kk.x, kk.y, h0, l0, bad);
    split_h2(kk.z, kk.w, h1, l1, bad);
    *reinterpret_cast<uint2*>(k_hi + key * kKLd + c4) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(k_lo + key * kKLd + c4) = make_uint2(l0, l1);
    split_h2(vv.x, vv.y, h0, l0, bad);
    split_h2(vv.z, vv.w, h1, l1, bad);
    *reinterpret_cast<uint2*>(v_hi + key * kKLd + c4) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(v_lo + key * kKLd + c4) = make_uint2(l0, l1);
  }
  __syncthreads();
  const int nrows = (a.ragged ? a.lay.P - t0 : a.nblocks) * rpq;
  const int64_t q_ld = a.q_ld ? a.q_ld : inner;
  // q / rpq by multiplication: exact while q * rpq < 2^32 (q < 32 positions x rpq rows, rpq <= 2048 beams)
  const uint32_t rpq_magic = (uint32_t)(0x100000000ull / (uint32_t)rpq) + ((0x100000000ull % (uint32_t)rpq) ? 1u : 0u);
  auto global_row = [&](int q) {
    const int blk = rpq == 1 ? q : (int)__umulhi((uint32_t)q, rpq_magic);
    return (a.ragged ? (int64_t)a.lay.off[t0 + blk] : (int64_t)blk * a.block_rows) + (int64_t)b * rpq + (q - blk * rpq);
  };
  const uint32_t* kh32 = reinterpret_cast<const uint32_t*>(k_hi);
  const uint32_t* kl32 = reinterpret_cast<const uint32_t*>(k_lo);
  const int vrow = (lane & 7) + 8 * ((lane >> 3) & 1), vcol = 8 * (lane >> 4);   // ldmatrix row of this lane
  // gridDim.y CTAs share one (query, head): each restages the <= 32 keys and takes every gridDim.y-th group of 4 tiles
  for (int tile = blockIdx.y * 4 + warp; tile * 16 < nrows; tile += 4 * gridDim.y) {
    const int q0 = tile * 16 + g, q1 = q0 + 8;
    const int64_t row0g = global_row(min(q0, nrows - 1)), row1g = global_row(min(q1, nrows - 1));   // (divisions: once)
    const float* qr0 = a.q + row0g * q_ld + h * 64 + 2 * t;
    const float* qr1 = a.q + row1g * q_ld + h * 64 + 2 * t;
    float2 qa[4][4];                                                // A fragments (raw fp32 pairs) of the 4 k-steps
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      qa[kk][0] = __ldg(reinterpret_cast<const float2*>(qr0 + kk * 16));
      qa[kk][1] = __ldg(reinterpret_cast<const float2*>(qr1 + kk * 16));
      qa[kk][2] = __ldg(reinterpret_cast<const float2*>(qr0 + kk * 16 + 8));
      qa[kk][3] = __ldg(reinterpret_cast<const float2*>(qr1 + kk * 16 + 8));
    }
    float sacc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int u = 0; u < 4; ++u) sacc[nt][u] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t ah[4], al[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) split_h2(qa[kk][u].x, qa[kk][u].y, ah[u], al[u], bad);
#pragma unroll
      // term-major over the four key tiles: consecutive MMAs write different accumulators (the three products of
      // one accumulator issued back to back wait for each other's result); per accumulator the order is unchanged
      uint32_t bh0[4], bh1[4], bl0[4], bl1[4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int o0 = (nt * 8 + g) * (kKLd / 2) + kk * 8 + t;      // words: K[key nt*8+g][dims kk*16 + 2t, +1]
        bh0[nt] = kh32[o0]; bh1[nt] = kh32[o0 + 4]; bl0[nt] = kl32[o0]; bl1[nt] = kl32[o0 + 4];
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_f16(sacc[nt], al, bh0[nt], bh1[nt]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_f16(sacc[nt], ah, bl0[nt], bl1[nt]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) mma_f16(sacc[nt], ah, bh0[nt], bh1[nt]);
    }
    if (a.rel_bias != nullptr) {
      // encoder self-attention through this kernel (the S rows of a sequence are the "beams" of its query slot):
      // relative position bias rel_bias[h][(key - query position) + rel_S - 1] on the score fragments
      const float* rb = a.rel_bias + (int64_t)h * (2 * a.rel_S - 1) + (a.rel_S - 1);
      const int c0 = min(q0, nrows - 1), c1 = min(q1, nrows - 1);
      const int p0 = c0 - (rpq == 1 ? c0 : (int)__umulhi((uint32_t)c0, rpq_magic)) * rpq;     // position in the sequence
      const int p1 = c1 - (rpq == 1 ? c1 : (int)__umulhi((uint32_t)c1, rpq_magic)) * rpq;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int key = min(nt * 8 + 2 * t + w, S - 1);
          sacc[nt][w] += __ldg(rb + (key - p0));
          sacc[nt][2 + w] += __ldg(rb + (key - p1));
        }
    }
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const bool on = (valid >> (nt * 8 + 2 * t + w)) & 1u;
        sacc[nt][w] = on ? sacc[nt][w] : -INFINITY;
        sacc[nt][2 + w] = on ? sacc[nt][2 + w] : -INFINITY;
        mx0 = fmaxf(mx0, sacc[nt][w]);
        mx1 = fmaxf(mx1, sacc[nt][2 + w]);
      }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        sacc[nt][w] = mx0 == -INFINITY ? 0.f : expf(sacc[nt][w] - mx0);
        sacc[nt][2 + w] = mx1 == -INFINITY ? 0.f : expf(sacc[nt][2 + w] - mx1);
        sum0 += sacc[nt][w];
        sum1 += sacc[nt][2 + w];
      }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // O = P V: the C fragments of key tiles 2ks and 2ks+1 are exactly the A fragment of 16-key step ks
    float oacc[8][4];
#pragma unroll
    for (int nd = 0; nd < 8; ++nd)
#pragma unroll
      for (int u = 0; u < 4; ++u) oacc[nd][u] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t ph[4], pl[4];
      bool nb_ = false;                                             // probabilities are in [0, 1]
      split_h2(sacc[2 * ks][0], sacc[2 * ks][1], ph[0], pl[0], nb_);
      split_h2(sacc[2 * ks][2], sacc[2 * ks][3], ph[1], pl[1], nb_);
      split_h2(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1], ph[2], pl[2], nb_);
      split_h2(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3], ph[3], pl[3], nb_);
#pragma unroll
      for (int nd = 0; nd < 8; nd += 4) {                        // four dim tiles at a time, term-major
        uint32_t bh[2][4], bl[2][4];                               // [pair][b0, b1 of dim tile 2*pair, b0, b1 of the next]
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
          ldmatrix_x4_trans(bh[pr], v_hi + (ks * 16 + vrow) * kKLd + (nd + 2 * pr) * 8 + vcol);
          ldmatrix_x4_trans(bl[pr], v_lo + (ks * 16 + vrow) * kKLd + (nd + 2 * pr) * 8 + vcol);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(oacc[nd + j], pl, bh[j >> 1][2 * (j & 1)], bh[j >> 1][2 * (j & 1) + 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(oacc[nd + j], ph, bl[j >> 1][2 * (j & 1)], bl[j >> 1][2 * (j & 1) + 1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_f16(oacc[nd + j], ph, bh[j >> 1][2 * (j & 1)], bh[j >> 1][2 * (j & 1) + 1]);
      }
    }
    const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f, inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;
    __half* ctx_hi = static_cast<__half*>(ctx.base) + h * 64;
    store_row_planes_f16(ctx_hi + row0g * inner, ctx.plane, oacc, 0, inv0, q0 < nrows, t, bad);
    store_row_planes_f16(ctx_hi + row1g * inner, ctx.plane, oacc, 2, inv1, q1 < nrows, t, bad);
  }
  if (bad && ctx.overflow) *ctx.overflow = 1;
  pdl_trigger();
}

// Forced tail self-attention on the warp-level tensor cores (fp16x3 mode): one warp per (row, head); the lineage's
// K/V (cached prefix through the ancestry table + the T new positions) are split into fp16 planes in the warp's
// shared memory once (V transposed), then the T queries run as two 16-row m16n8k16 tiles with the causal mask and
// the relative position bias applied to the score fragments.
constexpr int kTailMmaWarpBytes = 4 * 32 * kKLd * 2;          // K hi | K lo | V hi | V lo

// The T queries of one (frozen row, head) task as 16-row m16n8k16 tiles against the staged fp16 K/V planes of its
// lineage: tile row q = query index (position t + q); causal mask and relative position bias on the score fragments.
// Q fragments (fp16 hi / lo words of the 4 k-steps) of tile `tile` of a task, straight from the planes
__device__ __forceinline__ void tail_load_q_planes(const TailAttnArgs& a, int rp, int h, int t, int T, int tile,
                                                   uint32_t (&qh)[4][4], uint32_t (&ql)[4][4]) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int inner = a.H * 64;
  const int q0 = tile * 16 + g, q1 = q0 + 8;
  const __half* qh0 = a.qkv_hi + ((int64_t)a.lay.off[t + min(q0, T - 1)] + rp) * 3 * inner + h * 64 + 2 * t4;
  const __half* qh1 = a.qkv_hi + ((int64_t)a.lay.off[t + min(q1, T - 1)] + rp) * 3 * inner + h * 64 + 2 * t4;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    qh[kk][0] = *reinterpret_cast<const uint32_t*>(qh0 + kk * 16);
    qh[kk][1] = *reinterpret_cast<const uint32_t*>(qh1 + kk * 16);
    qh[kk][2] = *reinterpret_cast<const uint32_t*>(qh0 + kk * 16 + 8);
    qh[kk][3] = *reinterpret_cast<const uint32_t*>(qh1 + kk * 16 + 8);
    ql[kk][0] = *reinterpret_cast<const uint32_t*>(qh0 + a.qkv_plane + kk * 16);
    ql[kk][1] = *reinterpret_cast<const uint32_t*>(qh1 + a.qkv_plane + kk * 16);
    ql[kk][2] = *reinterpret_cast<const uint32_t*>(qh0 + a.qkv_plane + kk * 16 + 8);
    ql[kk][3] = *reinterpret_cast<const uint32_t*>(qh1 + a.qkv_plane + kk * 16 + 8);
  }
}

// same from fp32 q rows, split here (precision modes / paths without planes)
__device__ __forceinline__ void tail_load_q_f32(const TailAttnArgs& a, int rp, int h, int t, int T, int tile,
                                                uint32_t (&qh)[4][4], uint32_t (&ql)[4][4], bool& bad) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int inner = a.H * 64;
  const int q0 = tile * 16 + g, q1 = q0 + 8;
  const float* qr0 = a.qkv + ((int64_t)a.lay.off[t + min(q0, T - 1)] + rp) * 3 * inner + h * 64 + 2 * t4;
  const float* qr1 = a.qkv + ((int64_t)a.lay.off[t + min(q1, T - 1)] + rp) * 3 * inner + h * 64 + 2 * t4;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    const float2 x0 = *reinterpret_cast<const float2*>(qr0 + kk * 16), x1 = *reinterpret_cast<const float2*>(qr1 + kk * 16);
    const float2 x2 = *reinterpret_cast<const float2*>(qr0 + kk * 16 + 8);
    const float2 x3 = *reinterpret_cast<const float2*>(qr1 + kk * 16 + 8);
    split_h2(x0.x, x0.y, qh[kk][0], ql[kk][0], bad);
    split_h2(x1.x, x1.y, qh[kk][1], ql[kk][1], bad);
    split_h2(x2.x, x2.y, qh[kk][2], ql[kk][2], bad);
    split_h2(x3.x, x3.y, qh[kk][3], ql[kk][3], bad);
  }
}

// One 16-row tile of a (frozen row, head) task against the staged fp16 K/V planes of its lineage: tile row q = query
// index (position t + q); S = Q K^T and O = softmax(S) V on m16n8k16 with the 3-product split, causal mask and
// relative position bias on the score fragments.
__device__ __forceinline__ void tail_mma16_tile(const TailAttnArgs& a, const ActOut& ctx, const __half* k_hi,
                                                const __half* k_lo, const __half* v_hi, const __half* v_lo, int rp,
                                                int h, int t, int T, int tile, const uint32_t (&qh)[4][4],
                                                const uint32_t (&ql)[4][4], bool& bad) {
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const uint32_t* kh32 = reinterpret_cast<const uint32_t*>(k_hi);
  const uint32_t* kl32 = reinterpret_cast<const uint32_t*>(k_lo);
  const int vrow = (lane & 7) + 8 * ((lane >> 3) & 1), vcol = 8 * (lane >> 4);   // ldmatrix row of this lane
  const int inner = a.H * 64, L = a.L;
  const int q0 = tile * 16 + g, q1 = q0 + 8;
  float sacc[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int u = 0; u < 4; ++u) sacc[nt][u] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
    // term-major over the four key tiles (see cross_attn_mma16_kernel): independent accumulators back to back
    uint32_t bh0[4], bh1[4], bl0[4], bl1[4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int o0 = (nt * 8 + g) * (kKLd / 2) + kk * 8 + t4;
      bh0[nt] = kh32[o0]; bh1[nt] = kh32[o0 + 4]; bl0[nt] = kl32[o0]; bl1[nt] = kl32[o0 + 4];
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16(sacc[nt], ql[kk], bh0[nt], bh1[nt]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16(sacc[nt], qh[kk], bl0[nt], bl1[nt]);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma_f16(sacc[nt], qh[kk], bh0[nt], bh1[nt]);
  }
  // causal mask + relative position bias: row q sits at position t + q and sees keys p <= t + q
  const int pq0 = t + q0, pq1 = t + q1;
  const float* bias_h = a.bias + h * L;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const int p = nt * 8 + 2 * t4 + w;
      sacc[nt][w] = (p <= pq0 && q0 < T) ? sacc[nt][w] + __ldg(bias_h + (pq0 - p)) : -INFINITY;
      sacc[nt][2 + w] = (p <= pq1 && q1 < T) ? sacc[nt][2 + w] + __ldg(bias_h + (pq1 - p)) : -INFINITY;
      mx0 = fmaxf(mx0, sacc[nt][w]);
      mx1 = fmaxf(mx1, sacc[nt][2 + w]);
    }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      sacc[nt][w] = mx0 == -INFINITY ? 0.f : expf(sacc[nt][w] - mx0);
      sacc[nt][2 + w] = mx1 == -INFINITY ? 0.f : expf(sacc[nt][2 + w] - mx1);
      sum0 += sacc[nt][w];
      sum1 += sacc[nt][2 + w];
    }
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
  sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
  sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
  float oacc[8][4];
#pragma unroll
  for (int nd = 0; nd < 8; ++nd)
#pragma unroll
    for (int u = 0; u < 4; ++u) oacc[nd][u] = 0.f;
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    uint32_t ph[4], pl[4];
    bool nb_ = false;
    split_h2(sacc[2 * ks][0], sacc[2 * ks][1], ph[0], pl[0], nb_);
    split_h2(sacc[2 * ks][2], sacc[2 * ks][3], ph[1], pl[1], nb_);
    split_h2(sacc[2 * ks + 1][0], sacc[2 * ks + 1][1], ph[2], pl[2], nb_);
    split_h2(sacc[2 * ks + 1][2], sacc[2 * ks + 1][3], ph[3], pl[3], nb_);
#pragma unroll
    for (int nd = 0; nd < 8; nd += 4) {                          // four dim tiles at a time, term-major
      uint32_t bh[2][4], bl[2][4];
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        ldmatrix_x4_trans(bh[pr], v_hi + (ks * 16 + vrow) * kKLd + (nd + 2 * pr) * 8 + vcol);
        ldmatrix_x4_trans(bl[pr], v_lo + (ks * 16 + vrow) * kKLd + (nd + 2 * pr) * 8 + vcol);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_f16(oacc[nd + j], pl, bh[j >> 1][2 * (j & 1)], bh[j >> 1][2 * (j & 1) + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_f16(oacc[nd + j], ph, bl[j >> 1][2 * (j & 1)], bl[j >> 1][2 * (j & 1) + 1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_f16(oacc[nd + j], ph, bh[j >> 1][2 * (j & 1)], bh[j >> 1][2 * (j & 1) + 1]);
    }
  }
  const float inv0 = sum0 > 0.f ? 1.0f / sum0 : 0.f, inv1 = sum1 > 0.f ? 1.0f / sum1 : 0.f;
  __half* ctx_hi = static_cast<__half*>(ctx.base) + h * 64;
  store_row_planes_f16(ctx_hi + ((int64_t)a.lay.off[t + min(q0, T - 1)] + rp) * inner, ctx.plane, oacc, 0, inv0, q0 < T,
                       t4, bad);
  store_row_planes_f16(ctx_hi + ((int64_t)a.lay.off[t + min(q1, T - 1)] + rp) * inner, ctx.plane, oacc, 2, inv1, q1 < T,
                       t4, bad);
}

// all tiles of a task; with planes the next tile's Q words are requested before the current tile computes
__device__ __forceinline__ void tail_mma16_tiles(const TailAttnArgs& a, const ActOut& ctx, const __half* k_hi,
                                                 const __half* k_lo, const __half* v_hi, const __half* v_lo, int rp,
                                                 int h, int t, int T, bool& bad) {
  uint32_t qh[4][4], ql[4][4];
  const bool qplanes = a.qkv_hi != nullptr;
  uint32_t qhn[4][4], qln[4][4];
  if (qplanes) tail_load_q_planes(a, rp, h, t, T, 0, qh, ql);
  else tail_load_q_f32(a, rp, h, t, T, 0, qh, ql, bad);
#pragma unroll 1
  for (int tile = 0; tile * 16 < T; ++tile) {
    const bool more = (tile + 1) * 16 < T;
    if (more) {
      if (qplanes) tail_load_q_planes(a, rp, h, t, T, tile + 1, qhn, qln);
      else tail_load_q_f32(a, rp, h, t, T, tile + 1, qhn, qln, bad);
    }
    tail_mma16_tile(a, ctx, k_hi, k_lo, v_hi, v_lo, rp, h, t, T, tile, qh, ql, bad);
    if (more) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int u = 0; u < 4; ++u) { qh[kk][u] = qhn[kk][u]; ql[kk][u] = qln[kk][u]; }
    }
  }
}

__device__ __forceinline__ void cp_async16_any(void* dst_smem, const void* src) {
  const uint32_t d32 = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d32), "l"(src) : "memory");
}

// ncu on the round-1 version of this kernel (everything staged from fp32 and split here): 4200 warp instructions per
// task, ~60 % of them the fp32 -> fp16 hi/lo split of K, V and Q; 28 % issue utilisation at 12 warps per SM. With the
// qkv GEMM of the pass writing fp16 hi/lo planes (EPI_PLANES) the rows of the pass are copied plane-to-plane with
// cp.async (one 16-byte chunk per lane and position: 4 planes x 8 chunks) and only the few cached positions (< t)
// are still split here; Q fragments are read from the planes as they are.
__global__ void __launch_bounds__(kWarps * 32, 3) self_attn_tail_mma16_kernel(TailAttnArgs a, ActOut ctx) {
  extern __shared__ __align__(16) unsigned char tmsmem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int half = lane >> 4, l16 = lane & 15;
  __half* k_hi = reinterpret_cast<__half*>(tmsmem + warp * kTailMmaWarpBytes);
  __half* k_lo = k_hi + 32 * kKLd;
  __half* v_hi = k_lo + 32 * kKLd;
  __half* v_lo = v_hi + 32 * kKLd;
  const int inner = a.H * 64, L = a.L;
  const int P = a.lay.P;
  const int ntask = a.R * a.H;
  const bool planes = a.qkv_hi != nullptr;
  bool bad = false;
  // rows >= P are never written by a task: zero them once (their scores are masked, but 0 * garbage must not happen)
  for (int e = lane; e < (32 - P) * 8; e += 32) {
    const int p = P + (e >> 3), ch = e & 7;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    *reinterpret_cast<uint4*>(k_hi + p * kKLd + ch * 8) = z;
    *reinterpret_cast<uint4*>(k_lo + p * kKLd + ch * 8) = z;
    *reinterpret_cast<uint4*>(v_hi + p * kKLd + ch * 8) = z;
    *reinterpret_cast<uint4*>(v_lo + p * kKLd + ch * 8) = z;
  }
  // plane copy: lane -> (plane, 16-byte chunk of the 64-dim head row)
  const int pl = lane >> 3, ch = lane & 7;
  __half* dst_plane = (pl == 0 ? k_hi : (pl == 1 ? k_lo : (pl == 2 ? v_hi : v_lo))) + ch * 8;
  pdl_wait();
  for (int wid = blockIdx.x * kWarps + warp; wid < ntask; wid += gridDim.x * kWarps) {
    const int rp = wid / a.H, h = wid - rp * a.H;                   // frozen row (freeze order), head
    const int bq = a.fz_list[rp / a.nb];
    const int r = bq * a.nb + rp % a.nb;                            // original row id
    const int t = a.qstart ? a.qstart[bq] : 0, T = P - t;           // this row's pass: positions t..P-1
    __syncwarp();                                                   // the previous task's fragment reads are done
    if (planes) {
      const __half* src = a.qkv_hi + ((pl & 1) ? a.qkv_plane : 0) + ((pl >> 1) ? 2 * inner : inner) + h * 64 + ch * 8;
      for (int p = t; p < P; ++p)
        cp_async16_any(dst_plane + p * kKLd, src + ((int64_t)a.lay.off[p] + rp) * 3 * inner);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // ---- positions that arrive as fp32 (the cached prefix; everything when there are no planes): split here -------
    const int nconv = planes ? t : P;
    const int slot_l = lane < t ? lane * (int)a.row_cap + a.anc[(int64_t)r * L + lane] : -1;
#pragma unroll 2
    for (int it = 0; it < 16; ++it) {
      if (2 * it >= nconv) break;                                   // warp-uniform
      const int p = 2 * it + half;
      const int slot = __shfl_sync(0xffffffffu, slot_l, p);
      if (p < nconv) {
        float4 kk, vv;
        if (p < t) {
          kk = *reinterpret_cast<const float4*>(a.cache_k + (int64_t)slot * inner + h * 64 + l16 * 4);
          vv = *reinterpret_cast<const float4*>(a.cache_v + (int64_t)slot * inner + h * 64 + l16 * 4);
        } else {
          const float* row = a.qkv + ((int64_t)a.lay.off[p] + rp) * 3 * inner + h * 64 + l16 * 4;
          kk = *reinterpret_cast<const float4*>(row + inner);
          vv = *reinterpret_cast<const float4*>(row + 2 * inner);
        }
        uint32_t h0, l0, h1, l1;
        split_h2(kk.x, kk.y, h0, l0, bad);
        split_h2(kk.z, kk.w, h1, l1, bad);
        *reinterpret_cast<uint2*>(k_hi + p * kKLd + l16 * 4) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(k_lo + p * kKLd + l16 * 4) = make_uint2(l0, l1);
        split_h2(vv.x, vv.y, h0, l0, bad);
        split_h2(vv.z, vv.w, h1, l1, bad);
        *reinterpret_cast<uint2*>(v_hi + p * kKLd + l16 * 4) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(v_lo + p * kKLd + l16 * 4) = make_uint2(l0, l1);
      }
    }
    if (planes) asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    tail_mma16_tiles(a, ctx, k_hi, k_lo, v_hi, v_lo, rp, h, t, T, bad);
  }
  if (bad && ctx.overflow) *ctx.overflow = 1;
  pdl_trigger();
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
  // position blocks of the forced tail: walked inside the kernel when the staged K/V can stay resident (S <= 32),
  // one grid slice per block otherwise
  constexpr int kXWarps = x_warps<QS, XB>();
  const dim3 grid(ceil_div((int64_t)B * a.H, kXWarps), ceil_div(a.rows_per_query, XB),
                  a.S <= 32 ? 1 : (a.ragged ? a.lay.P : a.nblocks)),
      block(kXWarps * 32);
  constexpr size_t smem = (size_t)kXWarps * x_warp_floats<QS, XB>() * sizeof(float);
  static_assert(smem <= 227 * 1024, "cross-attention staging exceeds the shared memory of an SM");
  auto kern = cross_attn_warp_kernel<QS, XB>;
  static bool attr_set = false;
  if (!attr_set) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  return launch_pdl(kern, grid, block, smem, s, a, ctx);
}

// the tail kernels that read fp16 hi/lo planes written by the EPI_PLANES GEMM epilogue (fp16x3 mode only)
bool tail_self_attn_reads_planes(int mode) {
  const char* e = getenv("RB200_SELF_MMA");
  const char* p = getenv("RB200_TAIL_PLANES");
  return prec_is_fp16(mode) && !(e && e[0] == '0') && !(p && p[0] == '0');
}
int launch_self_attn_tail(const TailAttnArgs& a, ActOut ctx, cudaStream_t s) {
  RB_REQUIRE(a.lay.P >= 1 && a.lay.P <= RB_TAIL_MAX_L, "forced tail needs 1 <= P <= %d (P=%d)", RB_TAIL_MAX_L, a.lay.P);
  if (a.R == 0) return 0;
  static const bool use_mma = []() {
    const char* e = getenv("RB200_SELF_MMA");
    return !(e && e[0] == '0');
  }();
  const bool mma = use_mma && prec_is_fp16(ctx.mode);      // tensor-core kernel on fp16 planes; FFMA kernel otherwise
  static bool attr_set = false;
  static int sms = 0;
  if (!attr_set) {
    RB_CUDA(cudaFuncSetAttribute(self_attn_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)kWarps * kTailWarpFloats * sizeof(float))));
    RB_CUDA(cudaFuncSetAttribute(self_attn_tail_mma16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)kWarps * kTailMmaWarpBytes)));
    int dev = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    attr_set = true;
  }
  RB_REQUIRE(a.qkv_hi == nullptr || mma, "q | k | v planes are only read by the tensor-core tail kernel");
  const size_t smem = mma ? (size_t)kWarps * kTailMmaWarpBytes : (size_t)kWarps * kTailWarpFloats * sizeof(float);
  const int want = ceil_div((int64_t)a.R * a.H, kWarps);
  const int per_sm = mma ? 3 : 2;
  const dim3 grid(want < per_sm * sms ? want : per_sm * sms), block(kWarps * 32);
  if (mma) RB_CUDA(launch_pdl(self_attn_tail_mma16_kernel, grid, block, smem, s, a, ctx));
  else RB_CUDA(launch_pdl(self_attn_tail_kernel, grid, block, smem, s, a, ctx));
  launch_count()++;
  return 0;
}

bool launch_cross_attn_warp(const CrossAttnArgs& a, ActOut ctx, cudaStream_t s, int* status) {
  static const bool use_mma = []() {
    const char* e = getenv("RB200_XATTN_MMA");
    return !(e && e[0] == '0');
  }();
  // many rows per (query, head) against <= 32 keys (the forced tail): tensor-core kernel
  // (the exact fp32 mode keeps the FFMA kernel)
  // (the fp16 kernel also takes the encoder's self-attention: relative position bias on its score fragments;
  // RB200_ENC_MMA=0 keeps the encoder on the FFMA kernel)
  const char* enc_env = a.rel_bias != nullptr ? getenv("RB200_ENC_MMA") : nullptr;   // read per call (parity tests)
  const bool enc_mma = !(enc_env && enc_env[0] == '0');
  const bool bias_ok = a.rel_bias == nullptr || (enc_mma && prec_is_fp16(ctx.mode) && a.rel_S >= a.S);
  static const int min_rows = []() {               // rows per (query, head) from which the tensor-core kernel is used
    const char* e = getenv("RB200_XATTN_MINROWS");
    return e && atoi(e) > 0 ? atoi(e) : 32;
  }();
  if (use_mma && ctx.mode != 0 && a.S <= 32 && bias_ok &&
      (a.ragged || (int64_t)a.nblocks * a.rows_per_query >= min_rows)) {
    const int B = a.M / a.rows_per_query;
    // resident CTAs per SM the registers are budgeted for: 4 (128 registers, no spills) measured 212 us per tail
    // launch against 233 us at 3 (157 registers; 5 with 96 registers and a few spills: 208 us); RB200_XATTN_OCC=3
    // selects the round-1 build
    const char* oe = getenv("RB200_XATTN_OCC");
    const int occ = oe ? atoi(oe) : 4;
    // few (query, head) pairs with many rows each (the shipped --batch_size 1 --topk 1000 launch: 12 pairs of up to
    // 32 000 rows): split each pair's 64-row tile groups over gridDim.y CTAs until the grid fills the SMs 4 deep
    static const int sms = []() {
      int dev = 0, n = 148;
      if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
      return n;
    }();
    const int64_t max_rows = (int64_t)(a.ragged ? a.lay.P : a.nblocks) * a.rows_per_query;
    const int groups = (int)ceil_div(max_rows, (int64_t)64);
    int split = (int)ceil_div((int64_t)4 * sms, (int64_t)B * a.H);
    split = split < 1 ? 1 : (split > groups ? groups : split);
    const char* se = getenv("RB200_XATTN_SPLIT");
    if (se && atoi(se) > 0) split = atoi(se) > groups ? groups : atoi(se);
    const dim3 grid(B * a.H, split);
    const cudaError_t err = prec_is_fp16(ctx.mode)
                                ? (occ == 3 ? launch_pdl(cross_attn_mma16_kernel<3>, grid, dim3(128), 0, s, a, ctx)
                                            : launch_pdl(cross_attn_mma16_kernel<4>, grid, dim3(128), 0, s, a, ctx))
                                : launch_pdl(cross_attn_mma_kernel, grid, dim3(128), 0, s, a, ctx);
    *status = err == cudaSuccess ? 0 : fail(RB200_ERR_CUDA, "cross_attn_mma_kernel launch: %s", cudaGetErrorString(err));
    launch_count()++;
    return true;
  }
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
