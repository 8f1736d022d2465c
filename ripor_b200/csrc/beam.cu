// Prefix-constrained beam step and finalize on the GPU (integer trie walk + float64 candidate ranking).
//
// One CTA per query. Per step it restates, bit for bit on the float64 side, what the reference does in
// t5_pretrainer/tasks/generation.py:
//   :453-458  s = log_softmax(logits) if apply_log_softmax_for_scores else logits          (fp32)
//   :461      valid_mask from the DocID trie (here: walked in HBM, no host round trip)
//   :462      processed = s + (1 - valid_mask) * (-1e9)                                     (float64)
//   :463      cand = processed + beam_scores[:, None]                                       (float64)
//   :485-492  top-2*nb over the nb*V candidates of a query, parent = idx // V, token = idx % V
//   :496-507  BeamSearchScorer.process with eos=None keeps the first nb of them
//   :511      input_ids = cat(input_ids[beam_idx], tokens)
// plus the bookkeeping the KV-cached decoder needs instead of _reorder_cache (:517-518): a per-beam
// ancestry table saying which row holds the K/V of each earlier position.
// Ties between exactly equal candidates (only possible between -1e9-penalised duplicates) go to the lower
// flat index; torch.topk leaves that order unspecified.
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "beam.h"

using rb::TrieState;
using rb::TrieView;

namespace {

constexpr int kMaxPerThread = 128;   // candidates owned by one thread, tracked in a 128-bit taken mask

__device__ __forceinline__ bool cand_better(double va, int ia, double vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

struct StepArgs {
  TrieView tv;
  int t, nb, rpq, apply_ls, L, d_model;
  const float* logits;
  const double* sc_old;
  const TrieState* st_old;
  const int32_t* hist_old;
  const int32_t* anc_old;
  double* sc_new;
  TrieState* st_new;
  int32_t* hist_new;
  int32_t* anc_new;
  int32_t* parent_out;
  int32_t* token_out;
  const float* embed_table;
  float* next_x;
  int32_t* not_forced;
};

template <int THREADS, int KREG>   // KREG > 0: each thread keeps its (<= KREG) candidate values in registers
__global__ void __launch_bounds__(THREADS) beam_step_kernel(const StepArgs a) {
  constexpr int NW = THREADS / 32;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.nb, V = a.tv.V, words = a.tv.words, t = a.t;
  const int total = nb * V;
  rb::pdl_wait();

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* bs = reinterpret_cast<double*>(smem_raw);            // [nb]
  double* win_val = bs + nb;                                    // [nb]
  double* red_val = win_val + nb;                               // [32]
  int* red_idx = reinterpret_cast<int*>(red_val + 32);          // [32]
  int* win_idx = red_idx + 32;                                  // [nb]
  float* row_max = reinterpret_cast<float*>(win_idx + nb);      // [nb]
  float* row_log = row_max + nb;                                // [nb]
  uint32_t* allow = reinterpret_cast<uint32_t*>(row_log + nb);  // [nb, words]

  // ---- A. allowed-children bitmap of every beam (trie walk state -> V bits) -------------------------
  for (int i = warp; i < nb; i += NW) {
    const TrieState s = a.st_old[b * nb + i];
    uint32_t* bm = allow + i * words;
    const int n = s.hi - s.lo;
    for (int w = lane; w < words; w += 32)
      bm[w] = (n > 0 && s.node >= 0 && t < a.tv.L) ? a.tv.node_bitmap[(int64_t)s.node * words + w] : 0u;
    __syncwarp();
    if (n > 0 && s.node < 0 && t < a.tv.L && lane < n) {
      const int v = rb::trie_code(a.tv, (int64_t)s.lo + lane, t);
      atomicOr(&bm[v >> 5], 1u << (v & 31));
    }
    if (lane == 0) bs[i] = a.sc_old[b * nb + i];
    // ---- A2. optional fp32 log-softmax statistics of the beam's logits row -------------------------
    if (a.apply_ls) {
      const float* row = a.logits + (int64_t)(b * a.rpq + (a.rpq == 1 ? 0 : i)) * V;
      float m = -INFINITY;
      for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int v = lane; v < V; v += 32) sum += expf(row[v] - m);
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) { row_max[i] = m; row_log[i] = logf(sum); }
    }
  }
  __syncthreads();

  // ---- B. float64 candidate values; iterative arg-max for the best nb ------------------------------
  auto cand_val = [&](int c) -> double {
    const int i = c / V, v = c - i * V;
    float x = a.logits[(int64_t)(b * a.rpq + (a.rpq == 1 ? 0 : i)) * V + v];
    if (a.apply_ls) x = (x - row_max[i]) - row_log[i];
    const bool ok = (allow[i * words + (v >> 5)] >> (v & 31)) & 1u;
    const double processed = ok ? (double)x : (double)x + (-1e9);   // s + (1 - mask) * (-1e9)
    const double val = processed + bs[i];
    // a NaN logit (e.g. after an fp16x3 range overflow) must not derail the selection: rank it last, by index
    return val == val ? val : -1.7976931348623157e308;
  };
  uint32_t taken[kMaxPerThread / 32] = {0u, 0u, 0u, 0u};
  const int K = (total + THREADS - 1) / THREADS;
  // The thread that owns a winner looks for its next best candidate while everybody else waits at the barrier of the
  // next round: with the values cached in registers that is K compares instead of K x (division, logit load, bitmap
  // lookup, float64 add).
  constexpr int KR = KREG > 0 ? KREG : 1;
  const bool cached = KREG > 0 && K <= KREG;
  double cv[KR];
  if (cached) {
#pragma unroll
    for (int k = 0; k < KR; ++k) {
      const int c = tid + k * THREADS;
      cv[k] = (k < K && c < total) ? cand_val(c) : 0.0;
    }
  }
  double best_v;
  int best_c;
  auto rescan = [&]() {
    best_v = -INFINITY;
    best_c = INT_MAX;
    if (cached) {
#pragma unroll
      for (int k = 0; k < KR; ++k) {
        const int c = tid + k * THREADS;
        if (k < K && c < total && !((taken[0] >> k) & 1u) && cand_better(cv[k], c, best_v, best_c)) {
          best_v = cv[k];
          best_c = c;
        }
      }
      return;
    }
    for (int k = 0; k < K; ++k) {
      const int c = tid + k * THREADS;
      if (c < total && !((taken[k >> 5] >> (k & 31)) & 1u)) {
        const double v = cand_val(c);
        if (cand_better(v, c, best_v, best_c)) { best_v = v; best_c = c; }
      }
    }
  };
  rescan();
  for (int j = 0; j < nb; ++j) {
    double v = best_v;
    int c = best_c;
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_down_sync(0xffffffffu, v, o);
      const int oc = __shfl_down_sync(0xffffffffu, c, o);
      if (cand_better(ov, oc, v, c)) { v = ov; c = oc; }
    }
    if (lane == 0) { red_val[warp] = v; red_idx[warp] = c; }
    __syncthreads();
    if (warp == 0) {
      v = lane < NW ? red_val[lane] : -INFINITY;
      c = lane < NW ? red_idx[lane] : INT_MAX;
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oc = __shfl_down_sync(0xffffffffu, c, o);
        if (cand_better(ov, oc, v, c)) { v = ov; c = oc; }
      }
      if (lane == 0) { win_val[j] = v; win_idx[j] = c; }
    }
    __syncthreads();
    c = win_idx[j];
    if (c != INT_MAX && (c % THREADS) == tid) {
      const int k = c / THREADS;
      taken[k >> 5] |= 1u << (k & 31);
      rescan();
    }
  }

  rb::pdl_trigger();   // selection done; only the state write-back remains
  // ---- C. new beam state: scores, trie child, token history, KV ancestry, next decoder input --------
  // one warp per new beam: the child lookup (rb::trie_child, restated warp-wide) reads up to RB_TRIE_SMALL code rows
  // or the node's bitmap words, which a single thread would fetch one dependent load after the other
  for (int j = warp; j < nb; j += NW) {
    const int c = win_idx[j];
    const int i = c / V, v = c - i * V;
    const int r_new = b * nb + j;
    const TrieState s = a.st_old[b * nb + i];
    const int n = s.hi - s.lo;
    TrieState ns = rb::trie_dead();
    if (n > 0 && t < a.tv.L && v >= 0 && v < V) {
      if (s.node >= 0) {
        const uint32_t* bm = a.tv.node_bitmap + (int64_t)s.node * words;
        const int w = v >> 5;
        const uint32_t bit = 1u << (v & 31);
        int k = 0;
        uint32_t wv = 0;
        for (int w0 = 0; w0 <= w; w0 += 32) {                       // popcount rank of v among the node's children
          const int wi = w0 + lane;
          const uint32_t x = wi <= w ? bm[wi] : 0u;
          if (wi == w) wv = x;
          int part = wi < w ? __popc(x) : (wi == w ? __popc(x & (bit - 1u)) : 0);
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          k += part;
        }
        wv = __shfl_sync(0xffffffffu, wv, w & 31);
        if (wv & bit) {
          const int cidx = a.tv.node_child_ptr[s.node] + k;
          ns = TrieState{a.tv.child_lo[cidx], a.tv.child_lo[cidx + 1], a.tv.child_node[cidx], 0};
        }
      } else {
        int less = 0, leq = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {                        // n <= RB_TRIE_SMALL: one round
          const bool in = j0 + lane < n;
          const int cc = in ? rb::trie_code(a.tv, (int64_t)s.lo + j0 + lane, t) : INT_MAX;
          less += __popc(__ballot_sync(0xffffffffu, in && cc < v));
          leq += __popc(__ballot_sync(0xffffffffu, in && cc <= v));
        }
        if (leq != less) ns = TrieState{s.lo + less, s.lo + leq, -1, 0};
      }
    }
    if (lane == 0) {
      a.sc_new[r_new] = win_val[j];
      a.parent_out[r_new] = i;
      a.token_out[r_new] = v;
      a.st_new[r_new] = ns;
      if (ns.hi - ns.lo != 1) atomicAdd(a.not_forced, 1);   // not (yet) a single leaf: the tail cannot be forced
    }
  }
  const int L = a.L;
  for (int e = tid; e < nb * L; e += THREADS) {
    const int j = e / L, p = e - j * L;
    const int c = win_idx[j];
    const int i = c / V, v = c - i * V;
    const int src = b * nb + i, dst = b * nb + j;
    a.hist_new[dst * L + p] = p < t ? a.hist_old[src * L + p] : (p == t ? v : 0);
    int anc;
    if (p < t) anc = a.anc_old[src * L + p];
    else if (p == t) anc = (a.rpq == 1) ? b : src;    // the row that ran position t for this lineage
    else if (p == t + 1) anc = dst;                   // next step attends to itself at position t+1
    else anc = 0;
    a.anc_new[dst * L + p] = anc;
  }
  if (a.embed_table != nullptr) {
    const int d = a.d_model;
    for (int e = tid; e < nb * d; e += THREADS) {
      const int j = e / d, col = e - j * d;
      const int v = win_idx[j] % V;
      a.next_x[(int64_t)(b * nb + j) * d + col] = a.embed_table[(int64_t)v * d + col];
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-per-query formulation of the same step (nb <= kWarpNb): no block barriers. Every lane keeps the best nb of its
// own candidates (flat index = lane + 32 k) in a small sorted list; the nb winners are then popped with warp-wide
// arg-max rounds over the list heads. The global top-nb under (value desc, flat index asc) is contained in the union
// of the per-lane top-nb lists, so the result is identical to the block kernel's (and to the reference's top-k).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpNb = 16;          // beams per query this kernel handles
constexpr int kWarpQ = 4;            // queries (warps) per CTA
constexpr int kListLd = kWarpNb + 1; // padded per-lane list stride

__global__ void __launch_bounds__(kWarpQ * 32) beam_step_warp_kernel(const StepArgs a, int batch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kWarpQ + warp;
  const int nb = a.nb, V = a.tv.V, words = a.tv.words, t = a.t, L = a.L;
  const int total = nb * V;
  rb::pdl_wait();
  if (b >= batch) return;

  extern __shared__ __align__(16) unsigned char wsmem_raw[];
  // per warp: list values [32][kListLd] f64 | list indices [32][kListLd] i32 | bs[kWarpNb] f64 | win_val[kWarpNb] f64 |
  //           win_idx[kWarpNb] | row_max, row_log [kWarpNb] f32 | allow[kWarpNb * words]
  const size_t per_warp = 32 * kListLd * 12 + kWarpNb * (8 + 8 + 4 + 4 + 4) + (size_t)kWarpNb * words * 4;
  unsigned char* base = wsmem_raw + warp * ((per_warp + 15) & ~(size_t)15);
  double* list_v = reinterpret_cast<double*>(base);
  double* bs = list_v + 32 * kListLd;
  double* win_val = bs + kWarpNb;
  int* list_c = reinterpret_cast<int*>(win_val + kWarpNb);
  int* win_idx = list_c + 32 * kListLd;
  float* row_max = reinterpret_cast<float*>(win_idx + kWarpNb);
  float* row_log = row_max + kWarpNb;
  uint32_t* allow = reinterpret_cast<uint32_t*>(row_log + kWarpNb);

  // ---- A. allowed-children bitmaps, beam scores, optional log-softmax statistics ---------------------------------
  // The trie tables are far larger than L2 (codes: 280 MB at 8.8 M documents), so every lookup is a DRAM round trip:
  // the loads of all nb beams are issued together instead of one beam after the other.
  const TrieState my = lane < nb ? a.st_old[b * nb + lane] : rb::trie_dead();   // lane i = beam i
  if (lane < nb) bs[lane] = a.sc_old[b * nb + lane];
  {
    const int nbw = nb * words;
    for (int idx0 = 0; idx0 < nbw; idx0 += 32) {                    // explicit nodes: (beam, word) pairs over the lanes
      const int idx = min(idx0 + lane, nbw - 1);
      const int i = idx / words, w = idx - i * words;
      const int node_i = __shfl_sync(0xffffffffu, my.node, i);
      const int n_i = __shfl_sync(0xffffffffu, my.hi - my.lo, i);
      if (idx0 + lane < nbw)
        allow[idx] = (n_i > 0 && node_i >= 0 && t < a.tv.L) ? a.tv.node_bitmap[(int64_t)node_i * words + w] : 0u;
    }
    __syncwarp();
    int code[kWarpNb];                                              // implicit ranges: column t of their <= 32 rows
#pragma unroll
    for (int i = 0; i < kWarpNb; ++i) {
      const int node_i = __shfl_sync(0xffffffffu, my.node, i & 31);
      const int lo_i = __shfl_sync(0xffffffffu, my.lo, i & 31);
      const int n_i = __shfl_sync(0xffffffffu, my.hi - my.lo, i & 31);
      code[i] = (i < nb && n_i > 0 && node_i < 0 && t < a.tv.L && lane < n_i)
                    ? rb::trie_code(a.tv, (int64_t)lo_i + lane, t) : -1;
    }
#pragma unroll
    for (int i = 0; i < kWarpNb; ++i)
      if (code[i] >= 0) atomicOr(&allow[i * words + (code[i] >> 5)], 1u << (code[i] & 31));
  }
  if (a.apply_ls) {
    for (int i = 0; i < nb; ++i) {
      const float* row = a.logits + (int64_t)(b * a.rpq + (a.rpq == 1 ? 0 : i)) * V;
      float m = -INFINITY;
      for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int v = lane; v < V; v += 32) sum += expf(row[v] - m);
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0) { row_max[i] = m; row_log[i] = logf(sum); }
    }
  }
  __syncwarp();

  // ---- B1. every lane: sorted list of its best nb candidates -------------------------------------------------------
  double* lv = list_v + lane * kListLd;
  int* lc = list_c + lane * kListLd;
  int cnt = 0;
  {
    // walk (beam i, token v) with flat index c = i * V + v = lane + 32 k without divisions
    int i = 0, v = lane;
    while (v >= V) { v -= V; ++i; }
    for (int c = lane; c < total; c += 32) {
      float x = a.logits[(int64_t)(b * a.rpq + (a.rpq == 1 ? 0 : i)) * V + v];
      if (a.apply_ls) x = (x - row_max[i]) - row_log[i];
      const bool ok = (allow[i * words + (v >> 5)] >> (v & 31)) & 1u;
      const double processed = ok ? (double)x : (double)x + (-1e9);   // s + (1 - mask) * (-1e9)
      double val = processed + bs[i];
      val = val == val ? val : -1.7976931348623157e308;                // NaN logits rank last, by index
      if (cnt < nb || cand_better(val, c, lv[cnt - 1], lc[cnt - 1])) {
        int pos = cnt < nb ? cnt : nb - 1;                              // a full list drops its last entry
        while (pos > 0 && cand_better(val, c, lv[pos - 1], lc[pos - 1])) {
          lv[pos] = lv[pos - 1];
          lc[pos] = lc[pos - 1];
          --pos;
        }
        lv[pos] = val;
        lc[pos] = c;
        if (cnt < nb) ++cnt;
      }
      v += 32;
      while (v >= V) { v -= V; ++i; }
    }
  }
  // ---- B2. nb rounds of warp arg-max over the list heads ---------------------------------------------------------
  int hd = 0;
  for (int j = 0; j < nb; ++j) {
    double v = hd < cnt ? lv[hd] : -INFINITY;
    int c = hd < cnt ? lc[hd] : INT_MAX;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oc = __shfl_xor_sync(0xffffffffu, c, o);
      if (cand_better(ov, oc, v, c)) { v = ov; c = oc; }
    }
    if (hd < cnt && lc[hd] == c) ++hd;                                  // the owner pops its head
    if (lane == 0) { win_val[j] = v; win_idx[j] = c; }
  }
  __syncwarp();
  rb::pdl_trigger();

  // ---- C. new beam state: lane j = new beam j; again all dependent trie loads of the nb beams overlap ------------
  const int cj = lane < nb ? win_idx[lane] : 0;
  const int pj = cj / V, vj = cj - pj * V;                          // parent beam, token
  const TrieState sj = lane < nb ? a.st_old[b * nb + pj] : rb::trie_dead();
  const int nj = sj.hi - sj.lo;
  const bool live = lane < nb && nj > 0 && t < a.tv.L;
  TrieState ns = rb::trie_dead();
  if (live && sj.node >= 0) {                                       // explicit node: popcount rank, then the child slot
    const uint32_t* bm = a.tv.node_bitmap + (int64_t)sj.node * words;
    const int w = vj >> 5;
    const uint32_t bit = 1u << (vj & 31);
    int k = 0;
    for (int wi = 0; wi < w; ++wi) k += __popc(bm[wi]);
    const uint32_t wv = bm[w];
    if (wv & bit) {
      const int cidx = a.tv.node_child_ptr[sj.node] + k + __popc(wv & (bit - 1u));
      ns = TrieState{a.tv.child_lo[cidx], a.tv.child_lo[cidx + 1], a.tv.child_node[cidx], 0};
    }
  }
  {                                                                 // implicit ranges: warp-wide count per beam
    int code[kWarpNb];
#pragma unroll
    for (int j = 0; j < kWarpNb; ++j) {
      const int node_j = __shfl_sync(0xffffffffu, sj.node, j & 31);
      const int lo_j = __shfl_sync(0xffffffffu, sj.lo, j & 31);
      const int n_j = __shfl_sync(0xffffffffu, live ? nj : 0, j & 31);
      code[j] = (j < nb && n_j > 0 && node_j < 0 && lane < n_j) ? rb::trie_code(a.tv, (int64_t)lo_j + lane, t) : INT_MAX;
    }
#pragma unroll
    for (int j = 0; j < kWarpNb; ++j) {
      const int v_j = __shfl_sync(0xffffffffu, vj, j & 31);
      const int less = __popc(__ballot_sync(0xffffffffu, code[j] < v_j));
      const int leq = __popc(__ballot_sync(0xffffffffu, code[j] <= v_j));   // INT_MAX never counts (v_j < V)
      if (lane == j && live && sj.node < 0 && leq != less) ns = TrieState{sj.lo + less, sj.lo + leq, -1, 0};
    }
  }
  if (lane < nb) {
    const int r_new = b * nb + lane;
    a.sc_new[r_new] = win_val[lane];
    a.parent_out[r_new] = pj;
    a.token_out[r_new] = vj;
    a.st_new[r_new] = ns;
    if (ns.hi - ns.lo != 1) atomicAdd(a.not_forced, 1);
  }
  // token history and KV ancestry of the new beams: (beam j, position p) pairs over the lanes
  {
    const int tot = nb * L;
#pragma unroll 4
    for (int e = lane; e < tot; e += 32) {
      const int j = e / L, p = e - j * L;
      const int c = win_idx[j];
      const int i = c / V, v = c - i * V;
      const int src = b * nb + i, dst = b * nb + j;
      a.hist_new[dst * L + p] = p < t ? a.hist_old[src * L + p] : (p == t ? v : 0);
      int anc;
      if (p < t) anc = a.anc_old[src * L + p];
      else if (p == t) anc = (a.rpq == 1) ? b : src;
      else if (p == t + 1) anc = dst;
      else anc = 0;
      a.anc_new[dst * L + p] = anc;
    }
  }
  if (a.embed_table != nullptr) {                                    // next decoder input rows (d_model % 4 == 0)
    const int d4 = a.d_model >> 2, tot = nb * d4;
#pragma unroll 4
    for (int e = lane; e < tot; e += 32) {
      const int j = e / d4, col = e - j * d4;
      const int v = win_idx[j] % V;
      reinterpret_cast<float4*>(a.next_x + (int64_t)(b * nb + j) * a.d_model)[col] =
          reinterpret_cast<const float4*>(a.embed_table + (int64_t)v * a.d_model)[col];
    }
  }
}

__global__ void beam_reset_kernel(TrieView tv, int nb, int L, int batch, double* sc, TrieState* st, int32_t* hist,
                                  int32_t* anc) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= batch * nb) return;
  const int b = r / nb, i = r - b * nb;
  sc[r] = (i == 0) ? 0.0 : (double)(-1e9f);   // fp32 zeros with [:, 1:] = -1e9 (generation.py:418-420)
  st[r] = rb::trie_root(tv);
  for (int p = 0; p < L; ++p) { hist[r * L + p] = 0; anc[r * L + p] = (p == 0) ? b : 0; }
}

__global__ void beam_finalize_kernel(int nb, int L, int steps, int keep, double length_penalty,
                                     const double* __restrict__ sc, const TrieState* __restrict__ st,
                                     const int32_t* __restrict__ hist, int64_t* __restrict__ seqs,
                                     float* __restrict__ scores, int32_t* __restrict__ leaf) {
  // HF 4.17 BeamSearchScorer.finalize with eos=None: every beam becomes a hypothesis with
  // score / len**length_penalty (float64); sorted ascending (stable) and popped from the end.
  extern __shared__ double hs[];
  const int b = blockIdx.x;
  const double denom = pow((double)(steps + 1), length_penalty);
  for (int j = threadIdx.x; j < nb; j += blockDim.x) hs[j] = sc[b * nb + j] / denom;
  __syncthreads();
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    int rank = 0;
    for (int k = 0; k < nb; ++k) rank += (hs[k] > hs[j]) || (hs[k] == hs[j] && k > j);
    if (rank >= keep) continue;
    const int64_t o = (int64_t)b * keep + rank;
    const int r = b * nb + j;
    seqs[o * (steps + 1)] = 0;   // decoder_start_token_id
    for (int p = 0; p < steps; ++p) seqs[o * (steps + 1) + 1 + p] = hist[r * L + p];
    scores[o] = (float)hs[j];
    if (leaf) { leaf[o * 2] = st[r].lo; leaf[o * 2 + 1] = st[r].hi; }
  }
}

// forced tail: one CTA per beam row
__global__ void tail_prepare_kernel(TrieView tv, int t, int T, int R, int L, int d, const TrieState* __restrict__ st,
                                    int32_t* __restrict__ hist, const float* const* __restrict__ in_tabs,
                                    float* __restrict__ x) {
  const int r = blockIdx.x;
  const int64_t leaf = st[r].lo;
  for (int j = threadIdx.x; j < T; j += blockDim.x) hist[r * L + t + j] = rb::trie_code(tv, leaf, t + j);
  const int d4 = d >> 2;
  for (int j = 1; j < T; ++j) {
    const int tok = rb::trie_code(tv, leaf, t + j - 1);            // input of position t+j = token chosen at t+j-1
    const float4* src = reinterpret_cast<const float4*>(in_tabs[t + j - 1] + (int64_t)tok * d);
    float4* dst = reinterpret_cast<float4*>(x + ((int64_t)j * R + r) * d);
    for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = src[c];
  }
}

__global__ void tail_finish_kernel(TrieView tv, int t, int T, int R, int V, int apply_ls,
                                   const float* __restrict__ logits, const TrieState* __restrict__ st,
                                   double* __restrict__ sc) {
  // one warp per row; the float64 adds happen in step order, exactly as the step-by-step loop would do them
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= R) return;
  const int64_t leaf = st[r].lo;
  double s = sc[r];
  for (int j = 0; j < T; ++j) {
    const float* row = logits + ((int64_t)j * R + r) * V;
    float x = row[rb::trie_code(tv, leaf, t + j)];
    if (apply_ls) {
      float m = -INFINITY;
      for (int v = lane; v < V; v += 32) m = fmaxf(m, row[v]);
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      float sum = 0.f;
      for (int v = lane; v < V; v += 32) sum += expf(row[v] - m);
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      x = (x - m) - logf(sum);
    }
    s += (double)x;
  }
  if (lane == 0) sc[r] = s;
}

}  // namespace

namespace rb {

int launch_tail_prepare(rb200_beam* bm, const rb200_trie* trie, int T, const float* const* in_tabs_dev, float* x,
                        int d_model, cudaStream_t s) {
  const int R = bm->batch * bm->nb;
  tail_prepare_kernel<<<R, 128, 0, s>>>(trie->device_view(), bm->step, T, R, bm->L, d_model, bm->state[bm->cur],
                                        bm->hist[bm->cur], in_tabs_dev, x);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

int launch_tail_finish(rb200_beam* bm, const rb200_trie* trie, int T, const float* logits, int apply_log_softmax,
                       cudaStream_t s) {
  const int R = bm->batch * bm->nb;
  tail_finish_kernel<<<ceil_div(R, 4), 128, 0, s>>>(trie->device_view(), bm->step, T, R, bm->V, apply_log_softmax,
                                                    logits, bm->state[bm->cur], bm->scores[bm->cur]);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  bm->step += T;
  return 0;
}

}  // namespace rb

extern "C" {

int rb200_beam_create(int device, int max_batch, int num_beams, int L, int V, rb200_beam** out) {
  RB_REQUIRE(out, "null argument");
  RB_REQUIRE(max_batch >= 1 && num_beams >= 1 && L >= 1 && V >= 1, "need max_batch, num_beams, L, V >= 1");
  RB_REQUIRE((int64_t)num_beams * V <= 1024 * kMaxPerThread, "num_beams*V=%lld exceeds the beam kernel limit %d",
             (long long)num_beams * V, 1024 * kMaxPerThread);
  rb200_beam* bm = new (std::nothrow) rb200_beam();
  if (!bm) return rb::fail(RB200_ERR_NOMEM, "out of memory");
  bm->device = device; bm->max_batch = max_batch; bm->nb = num_beams; bm->L = L; bm->V = V;
  int prev = 0;
  RB_CUDA(cudaGetDevice(&prev));
  RB_CUDA(cudaSetDevice(device));
  const size_t R = (size_t)max_batch * num_beams;
  for (int h = 0; h < 2; ++h) {
    RB_CUDA(cudaMalloc(&bm->scores[h], R * sizeof(double)));
    RB_CUDA(cudaMalloc(&bm->state[h], R * sizeof(TrieState)));
    RB_CUDA(cudaMalloc(&bm->hist[h], R * L * sizeof(int32_t)));
    RB_CUDA(cudaMalloc(&bm->anc[h], R * L * sizeof(int32_t)));
  }
  RB_CUDA(cudaMalloc(&bm->not_forced, sizeof(int32_t)));
  RB_CUDA(cudaMalloc(&bm->parent, R * sizeof(int32_t)));
  RB_CUDA(cudaMalloc(&bm->token, R * sizeof(int32_t)));
  RB_CUDA(cudaSetDevice(prev));
  *out = bm;
  return 0;
}

int rb200_beam_free(rb200_beam* bm) {
  if (!bm) return 0;
  for (int h = 0; h < 2; ++h) {
    cudaFree(bm->scores[h]); cudaFree(bm->state[h]); cudaFree(bm->hist[h]); cudaFree(bm->anc[h]);
  }
  cudaFree(bm->parent); cudaFree(bm->token); cudaFree(bm->not_forced);
  delete bm;
  return 0;
}

int rb200_beam_reset(rb200_beam* bm, const rb200_trie* trie, int batch, void* stream) {
  RB_REQUIRE(bm && trie, "null argument");
  RB_REQUIRE(batch >= 1 && batch <= bm->max_batch, "batch %d outside [1, %d]", batch, bm->max_batch);
  RB_REQUIRE(trie->V == bm->V, "trie V=%d but beam state was created for V=%d", trie->V, bm->V);
  if (trie->device != bm->device)
    return rb::fail(RB200_ERR_STATE, "trie is on device %d, beam state on device %d: call rb200_trie_upload",
                    trie->device, bm->device);
  bm->batch = batch; bm->step = 0; bm->cur = 0;
  const int R = batch * bm->nb;
  beam_reset_kernel<<<rb::ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(
      trie->device_view(), bm->nb, bm->L, batch, bm->scores[0], bm->state[0], bm->hist[0], bm->anc[0]);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

int rb200_beam_step(rb200_beam* bm, const rb200_trie* trie, const float* logits, int rows_per_query,
                    int apply_log_softmax, const float* embed_table, float* next_x, int d_model, void* stream) {
  RB_REQUIRE(bm && trie && logits, "null argument");
  RB_REQUIRE(bm->batch >= 1, "rb200_beam_reset has not been called");
  RB_REQUIRE(rows_per_query == 1 || rows_per_query == bm->nb, "rows_per_query must be 1 or num_beams=%d", bm->nb);
  RB_REQUIRE(bm->step < bm->L, "already took L=%d steps", bm->L);
  RB_REQUIRE(bm->step < trie->L, "step %d exceeds the trie depth %d", bm->step, trie->L);
  RB_REQUIRE((embed_table == nullptr) == (next_x == nullptr), "embed table and next_x must be given together");
  StepArgs a;
  a.tv = trie->device_view();
  a.t = bm->step; a.nb = bm->nb; a.rpq = rows_per_query; a.apply_ls = apply_log_softmax; a.L = bm->L;
  a.d_model = d_model;
  a.logits = logits;
  const int o = bm->cur, n = bm->cur ^ 1;
  a.sc_old = bm->scores[o]; a.st_old = bm->state[o]; a.hist_old = bm->hist[o]; a.anc_old = bm->anc[o];
  a.sc_new = bm->scores[n]; a.st_new = bm->state[n]; a.hist_new = bm->hist[n]; a.anc_new = bm->anc[n];
  a.parent_out = bm->parent; a.token_out = bm->token;
  a.embed_table = embed_table; a.next_x = next_x;
  a.not_forced = bm->not_forced;
  RB_CUDA(cudaMemsetAsync(bm->not_forced, 0, sizeof(int32_t), (cudaStream_t)stream));
  const int nb = bm->nb;
  const size_t smem = (2 * nb + 32) * sizeof(double) + (32 + nb) * sizeof(int) + 2 * nb * sizeof(float) +
                      (size_t)nb * a.tv.words * sizeof(uint32_t);
  const int64_t total = (int64_t)nb * bm->V;
  static const bool force_cta = []() {
    const char* e = getenv("RB200_BEAM");
    return e && strcmp(e, "cta") == 0;
  }();
  const size_t per_warp = (32 * kListLd * 12 + kWarpNb * (8 + 8 + 4 + 4 + 4) + (size_t)kWarpNb * a.tv.words * 4 + 15) &
                          ~(size_t)15;
  // (very wide codebooks would push the per-warp bitmaps past the default 48 KB of dynamic shared memory)
  if (!force_cta && nb <= kWarpNb && d_model % 4 == 0 && kWarpQ * per_warp <= 48 * 1024) {
    RB_CUDA(rb::launch_pdl(beam_step_warp_kernel, dim3(rb::ceil_div(bm->batch, kWarpQ)), dim3(kWarpQ * 32),
                           kWarpQ * per_warp, (cudaStream_t)stream, a, bm->batch));
  } else if (total <= 256 * 32) {
    if (total <= 256 * 16)
      RB_CUDA(rb::launch_pdl(beam_step_kernel<256, 16>, dim3(bm->batch), dim3(256), smem, (cudaStream_t)stream, a));
    else
      RB_CUDA(rb::launch_pdl(beam_step_kernel<256, 32>, dim3(bm->batch), dim3(256), smem, (cudaStream_t)stream, a));
  } else {
    RB_CUDA(rb::launch_pdl(beam_step_kernel<1024, 0>, dim3(bm->batch), dim3(1024), smem, (cudaStream_t)stream, a));
  }
  rb::launch_count()++;
  bm->cur = n;
  bm->step += 1;
  return 0;
}

int rb200_beam_finalize(rb200_beam* bm, const rb200_trie* trie, int num_return, double length_penalty,
                        int64_t* sequences, float* scores, int32_t* leaf, void* stream) {
  RB_REQUIRE(bm && sequences && scores, "null argument");
  RB_REQUIRE(bm->batch >= 1 && bm->step >= 1, "no step has been taken");
  RB_REQUIRE(num_return >= 1 && num_return <= bm->nb,
             "`num_return_sequences` has to be smaller or equal to `num_beams`.");
  (void)trie;
  const int c = bm->cur;
  beam_finalize_kernel<<<bm->batch, 128, bm->nb * sizeof(double), (cudaStream_t)stream>>>(
      bm->nb, bm->L, bm->step, num_return, length_penalty, bm->scores[c], bm->state[c], bm->hist[c], sequences,
      scores, leaf);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

int rb200_beam_view(const rb200_beam* bm, int what, const void** ptr) {
  RB_REQUIRE(bm && ptr, "null argument");
  const int c = bm->cur;
  switch (what) {
    case 0: *ptr = bm->scores[c]; break;
    case 1: *ptr = bm->parent; break;
    case 2: *ptr = bm->token; break;
    case 3: *ptr = bm->hist[c]; break;
    case 4: *ptr = bm->anc[c]; break;
    case 5: *ptr = bm->state[c]; break;
    default: return rb::fail(RB200_ERR_INVALID, "unknown view %d", what);
  }
  return 0;
}

int rb200_beam_current_step(const rb200_beam* bm) { return bm ? bm->step : RB200_ERR_INVALID; }

}  // extern "C"
