// Prefix-constrained beam step and finalize on the GPU (integer trie walk + float64 candidate ranking).
//
// Per step and per query it restates, bit for bit on the float64 side, what the reference does in
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
// Ties between exactly equal candidates go to the lower flat index; torch.topk leaves that order unspecified.
//
// Three formulations of the same step: a warp per query (nb <= 16), a CTA per query with nb block-wide arg-max rounds
// (nb*V <= 131072), and a CTA per query with a radix select over the float64 keys (wide beams: the reference's shipped
// evaluation runs topk = 1000, full_scripts/full_evaluate_t5seq_aq_encoder.sh:191-199).
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "beam.h"

using rb::TrieState;
using rb::TrieView;

namespace {

constexpr int kMaxPerThread = 128;   // candidates owned by one thread of the arg-max kernel (128-bit taken mask)
constexpr double kNanRank = -1.7976931348623157e308;   // a NaN candidate ranks last, by index
// A query may be frozen only when every beam sits on a single leaf AND none carries the -1e9 penalty: only then is it
// certain that the beams' single valid children outrank every penalised candidate at each remaining step.
constexpr double kPenalised = -1e8;

__device__ __forceinline__ bool cand_better(double va, int ia, double vb, int ib) {
  return va > vb || (va == vb && ia < ib);
}

__device__ __forceinline__ unsigned long long f64_key(double v) {   // ascending order-preserving (-0.0 == +0.0)
  const unsigned long long b = (unsigned long long)__double_as_longlong(v + 0.0);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
  const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

struct StepArgs {
  TrieView tv;
  int t, nb, rpq, apply_ls, L, d_model;
  int nq;                       // queries of this launch (compact order)
  const int32_t* qlist;         // compact index -> query, nullptr = identity
  int allow_freeze;
  const float* logits;          // [nq * rpq, V] in compact order
  const double* sc_old;
  const TrieState* st_old;
  const int32_t* hist_old;
  const int32_t* anc_old;
  double* sc_new;
  TrieState* st_new;
  int32_t* hist_new;
  int32_t* anc_new;
  double* sc_fz;                // destinations of a query that freezes at this step
  TrieState* st_fz;
  int32_t* hist_fz;
  int32_t* anc_fz;
  int32_t* qstate;
  int32_t* parent_out;
  int32_t* token_out;
  const float* embed_table;
  float* next_x;                // rows at ORIGINAL row ids
};

// popcount rank of token v among the children of an explicit node + the child slot (warp-wide; all lanes get ns)
__device__ __forceinline__ TrieState child_explicit_warp(const TrieView& tv, const TrieState& s, int v, int lane) {
  const int words = tv.words;
  const uint32_t* bm = tv.node_bitmap + (int64_t)s.node * words;
  const int w = v >> 5;
  const uint32_t bit = 1u << (v & 31);
  int k = 0;
  uint32_t wv = 0;
  for (int w0 = 0; w0 <= w; w0 += 32) {
    const int wi = w0 + lane;
    const uint32_t x = wi <= w ? bm[wi] : 0u;
    if (wi == w) wv = x;
    int part = wi < w ? __popc(x) : (wi == w ? __popc(x & (bit - 1u)) : 0);
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    k += part;
  }
  wv = __shfl_sync(0xffffffffu, wv, w & 31);
  if (!(wv & bit)) return rb::trie_dead();
  const int cidx = tv.node_child_ptr[s.node] + k;
  return TrieState{tv.child_lo[cidx], tv.child_lo[cidx + 1], tv.child_node[cidx], 0};
}

// child of an implicit range (<= RB_TRIE_SMALL code rows) after token v at depth t (warp-wide)
__device__ __forceinline__ TrieState child_implicit_warp(const TrieView& tv, const TrieState& s, int t, int v, int lane) {
  const int n = s.hi - s.lo;
  int less = 0, leq = 0;
  for (int j0 = 0; j0 < n; j0 += 32) {
    const bool in = j0 + lane < n;
    const int cc = in ? rb::trie_code(tv, (int64_t)s.lo + j0 + lane, t) : INT_MAX;
    less += __popc(__ballot_sync(0xffffffffu, in && cc < v));
    leq += __popc(__ballot_sync(0xffffffffu, in && cc <= v));
  }
  if (leq == less) return rb::trie_dead();
  return TrieState{s.lo + less, s.lo + leq, -1, 0};
}

// state write-back shared by the two CTA kernels: win_val / win_idx hold the nb winners in rank order
template <int THREADS>
__device__ __forceinline__ void cta_write_back(const StepArgs& a, int b, const double* win_val, const int* win_idx,
                                               TrieState* ns_s, int* n_single) {
  constexpr int NW = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.nb, V = a.tv.V, t = a.t, L = a.L;
  if (tid == 0) *n_single = 0;
  __syncthreads();
  // one warp per new beam: the child lookup reads the node's bitmap words or up to RB_TRIE_SMALL code rows, which a
  // single thread would fetch one dependent load after the other
  for (int j = warp; j < nb; j += NW) {
    const int c = win_idx[j];
    const int i = c / V, v = c - i * V;
    const TrieState s = a.st_old[b * nb + i];
    const int n = s.hi - s.lo;
    TrieState ns = rb::trie_dead();
    if (n > 0 && t < a.tv.L && v >= 0 && v < V)
      ns = s.node >= 0 ? child_explicit_warp(a.tv, s, v, lane) : child_implicit_warp(a.tv, s, t, v, lane);
    if (lane == 0) {
      ns_s[j] = ns;
      if (ns.hi - ns.lo == 1 && win_val[j] > kPenalised) atomicAdd(n_single, 1);
    }
  }
  __syncthreads();
  const bool frozen = a.allow_freeze && *n_single == nb;     // every beam on a single leaf: the rest is determined
  double* sc_out = frozen ? a.sc_fz : a.sc_new;
  TrieState* st_out = frozen ? a.st_fz : a.st_new;
  int32_t* hist_out = frozen ? a.hist_fz : a.hist_new;
  int32_t* anc_out = frozen ? a.anc_fz : a.anc_new;
  if (tid == 0 && frozen) a.qstate[b] = t + 1;
  for (int j = tid; j < nb; j += THREADS) {
    const int c = win_idx[j];
    const int r_new = b * nb + j;
    sc_out[r_new] = win_val[j];
    a.parent_out[r_new] = c / V;
    a.token_out[r_new] = c % V;
    st_out[r_new] = ns_s[j];
  }
  for (int e = tid; e < nb * L; e += THREADS) {
    const int j = e / L, p = e - j * L;
    const int c = win_idx[j];
    const int i = c / V, v = c - i * V;
    const int src = b * nb + i, dst = b * nb + j;
    hist_out[dst * L + p] = p < t ? a.hist_old[src * L + p] : (p == t ? v : 0);
    int anc;
    if (p < t) anc = a.anc_old[src * L + p];
    else if (p == t) anc = (a.rpq == 1) ? b : src;    // the row that ran position t for this lineage
    else if (p == t + 1) anc = dst;                   // next step attends to itself at position t+1
    else anc = 0;
    anc_out[dst * L + p] = anc;
  }
  if (a.embed_table != nullptr) {
    // next decoder input rows: one warp per new beam, 16-byte copies (no per-element divisions: with 1000 beams the
    // element-wise version was ~30 % of this kernel, which runs on ONE SM per query)
    const int d = a.d_model;
    for (int j = warp; j < nb; j += NW) {
      const int v = win_idx[j] % V;
      const float* src = a.embed_table + (int64_t)v * d;
      float* dst = a.next_x + (int64_t)(b * nb + j) * d;
      if ((d & 3) == 0) {
        for (int c = lane; c < (d >> 2); c += 32)
          reinterpret_cast<float4*>(dst)[c] = reinterpret_cast<const float4*>(src)[c];
      } else {
        for (int c = lane; c < d; c += 32) dst[c] = src[c];
      }
    }
  }
}

// allowed-children bitmaps of every beam, beam scores and optional log-softmax statistics (CTA kernels)
template <int THREADS>
__device__ __forceinline__ void cta_prepare(const StepArgs& a, int b, int bc, double* bs, uint32_t* allow,
                                            float* row_max, float* row_log) {
  constexpr int NW = THREADS / 32;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.nb, V = a.tv.V, words = a.tv.words, t = a.t;
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
    if (a.apply_ls) {
      const float* row = a.logits + (int64_t)(bc * a.rpq + (a.rpq == 1 ? 0 : i)) * V;
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
}

// float64 candidate value of (beam i, token v) with logit x: s + (1 - mask) * (-1e9) + beam_score, NaN ranked last
__device__ __forceinline__ double cand_value_x(const StepArgs& a, float x, int i, int v, const double* bs,
                                               const uint32_t* allow, const float* row_max, const float* row_log) {
  if (a.apply_ls) x = (x - row_max[i]) - row_log[i];
  const bool ok = (allow[i * a.tv.words + (v >> 5)] >> (v & 31)) & 1u;
  const double processed = ok ? (double)x : (double)x + (-1e9);
  const double val = processed + bs[i];
  return val == val ? val : kNanRank;
}

__device__ __forceinline__ double cand_value(const StepArgs& a, int bc, int i, int v, const double* bs,
                                             const uint32_t* allow, const float* row_max, const float* row_log) {
  float x = a.logits[(int64_t)(bc * a.rpq + (a.rpq == 1 ? 0 : i)) * a.tv.V + v];
  if (a.apply_ls) x = (x - row_max[i]) - row_log[i];
  const bool ok = (allow[i * a.tv.words + (v >> 5)] >> (v & 31)) & 1u;
  const double processed = ok ? (double)x : (double)x + (-1e9);
  const double val = processed + bs[i];
  return val == val ? val : kNanRank;
}

template <int THREADS, int KREG>   // KREG > 0: each thread keeps its (<= KREG) candidate values in registers
__global__ void __launch_bounds__(THREADS) beam_step_kernel(const StepArgs a) {
  constexpr int NW = THREADS / 32;
  const int bc = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.nb, V = a.tv.V, words = a.tv.words;
  const int total = nb * V;
  rb::pdl_wait();
  const int b = a.qlist ? a.qlist[bc] : bc;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* bs = reinterpret_cast<double*>(smem_raw);            // [nb]
  double* win_val = bs + nb;                                    // [nb]
  double* red_val = win_val + nb;                               // [32]
  int* red_idx = reinterpret_cast<int*>(red_val + 32);          // [32]
  int* win_idx = red_idx + 32;                                  // [nb]
  float* row_max = reinterpret_cast<float*>(win_idx + nb);      // [nb]
  float* row_log = row_max + nb;                                // [nb]
  uint32_t* allow = reinterpret_cast<uint32_t*>(row_log + nb);  // [nb, words]
  TrieState* ns_s = reinterpret_cast<TrieState*>(allow + nb * words);   // [nb]
  __shared__ int n_single;

  cta_prepare<THREADS>(a, b, bc, bs, allow, row_max, row_log);

  // ---- float64 candidate values; iterative arg-max for the best nb ---------------------------------
  auto cand_val = [&](int c) -> double {
    const int i = c / V;
    return cand_value(a, bc, i, c - i * V, bs, allow, row_max, row_log);
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
  cta_write_back<THREADS>(a, b, win_val, win_idx, ns_s, &n_single);
}

// ---------------------------------------------------------------------------------------------------------------
// Wide beams (nb up to kSelMaxNb, nb*V up to a few million candidates per query): exact top-nb by a radix select.
// The float64 candidate values map to order-preserving 64-bit keys; 12-bit digits from the top narrow the bucket that
// holds the nb-th largest key until it fits the sort buffer, then everything above it plus the bucket is sorted by
// (value desc, flat index asc) with a bitonic network. A bucket of more than `cap` EXACTLY equal keys (e.g. step 0 of
// a search with nb > V: the nb-1 beams that start at -1e9 offer identical candidates, generation.py:418-420) is cut
// by a second radix select on the flat index, which reproduces "ties go to the lower flat index".
// Candidates are recomputed per pass from the logits (L2 resident) instead of being stored.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSelThreads = 1024;
constexpr int kSelBins = 4096;
constexpr int kSelMaxNb = 2048;
constexpr int kSelMaxSmem = 226 * 1024;   // dynamic part; the kernel also has a few hundred bytes of static shared memory


// histogram increment, aggregated over the lanes of the warp that hit the same bin: the first digit of the value
// select (sign + exponent) and the tie cut put hundreds of thousands of candidates into a handful of bins, and
// same-address shared-memory atomics serialise
__device__ __forceinline__ void hist_add(uint32_t* hist, unsigned bin) {
  const unsigned peers = __match_any_sync(__activemask(), bin);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
}

__global__ void __launch_bounds__(kSelThreads) beam_step_select_kernel(const StepArgs a, int pbuf) {
  const int bc = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nb = a.nb, V = a.tv.V, words = a.tv.words;
  const int total = nb * V;
  rb::pdl_wait();
  const int b = a.qlist ? a.qlist[bc] : bc;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* bs = reinterpret_cast<double*>(smem_raw);                       // [nb]
  double* win_val = bs + nb;                                               // [nb]
  unsigned long long* buf_key = reinterpret_cast<unsigned long long*>(win_val + nb);   // [pbuf]
  int* buf_idx = reinterpret_cast<int*>(buf_key + pbuf);                   // [pbuf]
  int* win_idx = buf_idx + pbuf;                                           // [nb]
  float* row_max = reinterpret_cast<float*>(win_idx + nb);                 // [nb]
  float* row_log = row_max + nb;                                           // [nb]
  uint32_t* hist = reinterpret_cast<uint32_t*>(row_log + nb);              // [kSelBins]
  uint32_t* allow = hist + kSelBins;                                       // [nb, words]
  TrieState* ns_s = reinterpret_cast<TrieState*>(allow + nb * words);      // [nb]
  __shared__ int n_single;
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_bucket, s_fill;
  __shared__ int s_part[32];

  cta_prepare<kSelThreads>(a, b, bc, bs, allow, row_max, row_log);

  // walk this thread's candidates c = tid + k * THREADS as (beam i, token v) without divisions
  const int step_i = kSelThreads / V, step_v = kSelThreads - step_i * V;
  auto for_each_cand = [&](auto&& fn) {
    constexpr int kBatch = 4;                  // logits of kBatch candidates in flight before any is consumed
    int i = tid / V, v = tid - i * V;
    for (int c0 = tid; c0 < total; c0 += kSelThreads * kBatch) {
      float xs[kBatch];
      int is[kBatch], vs[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        is[u] = i; vs[u] = v;
        xs[u] = (c0 + kSelThreads * u < total) ? a.logits[(int64_t)(bc * a.rpq + (a.rpq == 1 ? 0 : i)) * V + v] : 0.f;
        i += step_i;                           // c += kSelThreads without a division or a subtract loop
        v += step_v;
        if (v >= V) { v -= V; ++i; }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int c = c0 + kSelThreads * u;
        if (c < total) fn(c, f64_key(cand_value_x(a, xs[u], is[u], vs[u], bs, allow, row_max, row_log)));
      }
    }
  };
  // digit D of the histogram with count(> D) < need <= count(>= D) (from_top) or count(< D) < need <= count(<= D);
  // updates s_need (what is still needed inside the digit's bucket) and s_bucket (its population). Warp 0 only.
  auto pick_digit = [&](bool from_top) -> int {
    int D = 0;
    if (warp == 0) {
      const int seg = kSelBins / 32;
      int sum = 0;
      for (int k = 0; k < seg; ++k) sum += (int)hist[lane * seg + k];
      s_part[lane] = sum;
      __syncwarp();
      if (lane == 0) {
        const int need = s_need;
        int acc = 0, l = from_top ? 31 : 0;
        for (int n = 0; n < 32; ++n, l += from_top ? -1 : 1) {
          if (acc + s_part[l] >= need) break;
          acc += s_part[l];
        }
        l = l < 0 ? 0 : (l > 31 ? 31 : l);
        int k = from_top ? seg - 1 : 0;
        for (int n = 0; n < seg; ++n, k += from_top ? -1 : 1) {
          if (acc + (int)hist[l * seg + k] >= need) break;
          acc += (int)hist[l * seg + k];
        }
        k = k < 0 ? 0 : (k > seg - 1 ? seg - 1 : k);
        D = l * seg + k;
        s_need = need - acc;
        s_bucket = (int)hist[D];
      }
      D = __shfl_sync(0xffffffffu, D, 0);
    }
    return D;
  };

  // ---- 1. value radix select ------------------------------------------------------------------------------------
  const int cap = pbuf - nb;                 // room for the last bucket in the sort buffer
  if (tid == 0) { s_prefix = 0ull; s_need = nb < total ? nb : total; s_bucket = total; s_fill = 0; }
  __syncthreads();
  int bits_done = 0;
  while (bits_done < 64 && s_bucket > cap) {
    const int dbits = 64 - bits_done >= 12 ? 12 : 64 - bits_done;
    for (int k = tid; k < kSelBins; k += kSelThreads) hist[k] = 0u;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const int shift = 64 - bits_done - dbits;
    for_each_cand([&](int, unsigned long long key) {
      if (bits_done == 0 || (key >> (64 - bits_done)) == prefix)
        hist_add(hist, (unsigned)((key >> shift) & ((1u << dbits) - 1u)));
    });
    __syncthreads();
    const int D = pick_digit(true);
    if (tid == 0) s_prefix = (prefix << dbits) | (unsigned long long)D;
    bits_done += dbits;
    __syncthreads();
  }
  // ---- 2. collect: everything above the bucket, plus the bucket (or, for a bucket of > cap equal keys, its
  //         lowest flat indices) ---------------------------------------------------------------------------------
  const unsigned long long prefix = s_prefix;
  const bool tie_cut = s_bucket > cap;       // only possible with all 64 bits fixed: the bucket is one exact value
  auto in_bucket = [&](unsigned long long key) { return bits_done == 0 || (key >> (64 - bits_done)) == prefix; };
  auto above = [&](unsigned long long key) { return bits_done > 0 && (key >> (64 - bits_done)) > prefix; };
  int idx_hi = -1, idx_lim = INT_MAX;        // tie cut: take bucket members with flat index < idx_lim
  if (tie_cut) {
    // flat indices are < 2^24: two 12-bit passes from the top, smallest s_need of them
    for (int pass = 0; pass < 2; ++pass) {
      for (int k = tid; k < kSelBins; k += kSelThreads) hist[k] = 0u;
      __syncthreads();
      for_each_cand([&](int c, unsigned long long key) {
        if (key == prefix && (pass == 0 || (c >> 12) == idx_hi)) hist_add(hist, pass == 0 ? (c >> 12) : (c & 4095));
      });
      __syncthreads();
      const int D = pick_digit(false);
      if (tid == 0) s_part[0] = D;
      __syncthreads();
      if (pass == 0) idx_hi = s_part[0];
      else idx_lim = (idx_hi << 12) + s_part[0] + 1;
      __syncthreads();
    }
  }
  for_each_cand([&](int c, unsigned long long key) {
    if (above(key) || (in_bucket(key) && c < idx_lim)) {
      const int pos = atomicAdd(&s_fill, 1);
      if (pos < pbuf) { buf_key[pos] = key; buf_idx[pos] = c; }
    }
  });
  __syncthreads();
  const int filled = s_fill < pbuf ? s_fill : pbuf;
  for (int k = filled + tid; k < pbuf; k += kSelThreads) { buf_key[k] = 0ull; buf_idx[k] = INT_MAX; }
  __syncthreads();
  // ---- 3. bitonic sort of the buffer: (key desc, flat index asc) ---------------------------------------------------
  for (int k = 2; k <= pbuf; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < pbuf; i += kSelThreads) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long ki = buf_key[i], kp = buf_key[p];
          const int ci = buf_idx[i], cp = buf_idx[p];
          const bool i_first = ki > kp || (ki == kp && ci < cp);     // i belongs before p in the final order
          const bool up = (i & k) == 0;
          if (up ? !i_first : i_first) {
            buf_key[i] = kp; buf_key[p] = ki;
            buf_idx[i] = cp; buf_idx[p] = ci;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int j = tid; j < nb; j += kSelThreads) {
    win_val[j] = key_f64(buf_key[j]);
    win_idx[j] = buf_idx[j];
  }
  __syncthreads();
  rb::pdl_trigger();
  cta_write_back<kSelThreads>(a, b, win_val, win_idx, ns_s, &n_single);
}

// ---------------------------------------------------------------------------------------------------------------
// Warp-per-query formulation of the same step (nb <= kWarpNb): no block barriers. Every lane keeps the best nb of its
// own candidates (flat index = lane + 32 k) in a small sorted list; the nb winners are then popped with warp-wide
// arg-max rounds over the list heads. The global top-nb under (value desc, flat index asc) is contained in the union
// of the per-lane top-nb lists, so the result is identical to the block kernel's (and to the reference's top-k).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kWarpNb = 16;          // beams per query this kernel handles
constexpr int kWarpQ = 4;            // queries (warps) per CTA
// per warp: list values [32][nb + 1] f64 | bs[kWarpNb] f64 | win_val[kWarpNb] f64 | list indices [32][nb + 1] i32 |
//           win_idx[kWarpNb] | row_max, row_log, thr_ok, thr_pen [kWarpNb] f32 | allow[kWarpNb * words]
// (the per-lane lists are sized for the launch's nb, not for kWarpNb: shared memory is what bounds the resident warps)
__host__ __device__ constexpr size_t warp_step_smem(int words, int nb) {
  return ((size_t)32 * (nb + 1) * 12 + kWarpNb * (8 + 8 + 4 + 4 + 4 + 4 + 4) + (size_t)kWarpNb * words * 4 + 15) &
         ~(size_t)15;
}

template <int WORDS>   // codebook words known at compile time (8: V = 256, 32: V = 1024), 0 = read from the trie view
__global__ void __launch_bounds__(kWarpQ * 32) beam_step_warp_kernel(const StepArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bc = blockIdx.x * kWarpQ + warp;
  const int words = WORDS ? WORDS : a.tv.words, V = WORDS ? 32 * WORDS : a.tv.V;
  const int nb = a.nb, t = a.t, L = a.L;
  const int total = nb * V;
  rb::pdl_wait();
  if (bc >= a.nq) return;
  const int b = a.qlist ? a.qlist[bc] : bc;

  extern __shared__ __align__(16) unsigned char wsmem_raw[];
  const int list_ld = nb + 1;                  // padded per-lane list stride
  unsigned char* base = wsmem_raw + warp * warp_step_smem(words, nb);
  double* list_v = reinterpret_cast<double*>(base);
  double* bs = list_v + 32 * list_ld;
  double* win_val = bs + kWarpNb;
  int* list_c = reinterpret_cast<int*>(win_val + kWarpNb);
  int* win_idx = list_c + 32 * list_ld;
  float* row_max = reinterpret_cast<float*>(win_idx + kWarpNb);
  float* row_log = row_max + kWarpNb;
  float* thr_ok = row_log + kWarpNb;
  float* thr_pen = thr_ok + kWarpNb;
  uint32_t* allow = reinterpret_cast<uint32_t*>(thr_pen + kWarpNb);

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
    const unsigned implicit = __ballot_sync(0xffffffffu, lane < nb && my.hi - my.lo > 0 && my.node < 0 && t < a.tv.L);
#pragma unroll
    for (int i = 0; i < kWarpNb; ++i) {
      code[i] = -1;
      if ((implicit >> i) & 1u) {                                   // (warp-uniform: only beams on implicit ranges pay)
        const int lo_i = __shfl_sync(0xffffffffu, my.lo, i);
        const int n_i = __shfl_sync(0xffffffffu, my.hi - my.lo, i);
        if (lane < n_i) code[i] = rb::trie_code(a.tv, (int64_t)lo_i + lane, t);
      }
    }
#pragma unroll
    for (int i = 0; i < kWarpNb; ++i)
      if (code[i] >= 0) atomicOr(&allow[i * words + (code[i] >> 5)], 1u << (code[i] & 31));
  }
  if (a.apply_ls) {
    for (int i = 0; i < nb; ++i) {
      const float* row = a.logits + (int64_t)(bc * a.rpq + (a.rpq == 1 ? 0 : i)) * V;
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

  // ---- B1. every lane: sorted list of its best candidates ------------------------------------------------------------
  // Two passes over the lane's candidates (flat index c = i * V + v = lane + 32 k, walked without divisions; logits
  // requested kBatch at a time). Pass 1 finds a value of a real candidate per lane (its best); the nb-th largest of
  // these 32 values is a lower bound tau of the query's nb-th best candidate. Pass 2 inserts only candidates >= tau
  // into the lane's sorted list, so the data-dependent insertion - which a warp pays at the depth of its slowest lane
  // - runs for a few dozen candidates per query instead of for all nb * V of them.
  // V % 32 == 0 (every shipped codebook): lane owns token lane + 32 k of every beam and both passes stay in fp32 for
  // all but the surviving candidates. x -> value(x) = ((double)x [+ -1e9]) + beam score is non-decreasing in x inside
  // a (beam, allowed / penalised) class, so pass 1 needs one float64 evaluation per class (on the lane's largest x of
  // the class: the value of a real candidate), and pass 2 compares x with a per-class fp32 threshold rounded so that
  // x < threshold PROVES value(x) < tau (margin 2^-48 relative, far above the 2^-53 rounding of the three float64
  // operations; NaN / infinite thresholds fail the comparison and send the candidate to the exact evaluation).
  double* lv = list_v + lane * list_ld;
  int* lc = list_c + lane * list_ld;
  int cnt = 0;
  constexpr int kBatch = 8;
  const bool lane_tokens = WORDS ? true : (V & 31) == 0;
  auto exact_value = [&](float x, bool ok, double bsi) {
    const double processed = ok ? (double)x : (double)x + (-1e9);
    const double val = processed + bsi;
    return val == val ? val : kNanRank;
  };
  auto insert = [&](int c, double val) {
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
  };
  auto scan = [&](auto&& fn) {     // generic walk (V not a multiple of 32): float64 value of every candidate
    int i = 0, v = lane;
    while (v >= V) { v -= V; ++i; }
    for (int c0 = lane; c0 < total; c0 += 32 * kBatch) {
      float xs[kBatch];
      int is[kBatch], vs[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        is[u] = i; vs[u] = v;
        xs[u] = (c0 + 32 * u < total) ? a.logits[(int64_t)(bc * a.rpq + (a.rpq == 1 ? 0 : i)) * V + v] : 0.f;
        v += 32;
        while (v >= V) { v -= V; ++i; }
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int c = c0 + 32 * u;
        if (c < total) fn(c, cand_value_x(a, xs[u], is[u], vs[u], bs, allow, row_max, row_log));
      }
    }
  };
  double tau = -INFINITY;
  if (total > 32 * nb) {                       // (with few candidates per lane the filter saves nothing)
    double mine = -INFINITY;
    if (lane_tokens) {
      for (int i = 0; i < nb; ++i) {
        const float* row = a.logits + (int64_t)(bc * a.rpq + (a.rpq == 1 ? 0 : i)) * V + lane;
        const float rmax = a.apply_ls ? row_max[i] : 0.f, rlog = a.apply_ls ? row_log[i] : 0.f;
        float m_ok = -INFINITY, m_pen = -INFINITY;        // fmaxf drops NaNs: the maxima stay values of real candidates
        for (int k0 = 0; k0 < words; k0 += kBatch) {
          float xs[kBatch];
          uint32_t aw[kBatch];
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            xs[u] = (k0 + u < words) ? row[32 * (k0 + u)] : -INFINITY;
            aw[u] = (k0 + u < words) ? allow[i * words + k0 + u] : 0u;
          }
#pragma unroll
          for (int u = 0; u < kBatch; ++u) {
            float x = xs[u];
            if (a.apply_ls) x = (x - rmax) - rlog;
            const bool ok = (aw[u] >> lane) & 1u;
            m_ok = ok ? fmaxf(m_ok, x) : m_ok;
            m_pen = ok ? m_pen : fmaxf(m_pen, x);
          }
        }
        const double bsi = bs[i];
        if (m_ok > -INFINITY) {
          const double val = exact_value(m_ok, true, bsi);
          mine = val > mine ? val : mine;
        }
        if (m_pen > -INFINITY) {
          const double val = exact_value(m_pen, false, bsi);
          mine = val > mine ? val : mine;
        }
      }
    } else {
      scan([&](int, double val) { mine = val > mine ? val : mine; });
    }
    // nb-th largest of the lane values: nb warp-max rounds with removal, on order-preserving 64-bit keys (two 32-bit
    // redux.sync per round instead of five float64 shuffle + compare steps)
    unsigned long long key = f64_key(mine);
    unsigned long long tkey = 0ull;
    for (int j = 0; j < nb; ++j) {
      const unsigned mh = __reduce_max_sync(0xffffffffu, (unsigned)(key >> 32));
      const bool top = (unsigned)(key >> 32) == mh;
      const unsigned ml = __reduce_max_sync(0xffffffffu, top ? (unsigned)key : 0u);
      tkey = ((unsigned long long)mh << 32) | ml;
      const unsigned owners = __ballot_sync(0xffffffffu, key == tkey);
      if (lane == __ffs(owners) - 1) key = 0ull;               // remove ONE holder of the maximum (0 < every real key)
    }
    tau = tkey == 0ull ? -INFINITY : key_f64(tkey);            // (fewer than nb lanes held a candidate: no filter)
  }
  if (lane_tokens) {
    if (lane < nb) {                           // lane i: the two thresholds of beam i
      const double bsi = bs[lane];
      const double dlt = tau - bsi;
      const double mag = fabs(tau) + fabs(bsi);
      thr_ok[lane] = __double2float_rd(dlt - mag * 0x1p-48);
      thr_pen[lane] = __double2float_rd((dlt + 1e9) - (mag + 1e9) * 0x1p-47);
    }
    __syncwarp();
    for (int i = 0; i < nb; ++i) {
      const float* row = a.logits + (int64_t)(bc * a.rpq + (a.rpq == 1 ? 0 : i)) * V + lane;
      const float rmax = a.apply_ls ? row_max[i] : 0.f, rlog = a.apply_ls ? row_log[i] : 0.f;
      const float t_ok = thr_ok[i], t_pen = thr_pen[i];
      for (int k0 = 0; k0 < words; k0 += kBatch) {
        float xs[kBatch];
        uint32_t aw[kBatch];
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          xs[u] = (k0 + u < words) ? row[32 * (k0 + u)] : 0.f;
          aw[u] = (k0 + u < words) ? allow[i * words + k0 + u] : 0u;
        }
        uint32_t surv = 0u;                    // candidates the fp32 comparison could not rule out
#pragma unroll
        for (int u = 0; u < kBatch; ++u) {
          float x = xs[u];
          if (a.apply_ls) x = (x - rmax) - rlog;
          const bool ok = (aw[u] >> lane) & 1u;
          if (k0 + u < words && !(x < (ok ? t_ok : t_pen))) surv |= 1u << u;
        }
        if (surv != 0u) {
          const double bsi = bs[i];
          while (surv != 0u) {
            const int k = k0 + __ffs(surv) - 1;
            surv &= surv - 1u;
            float x = row[32 * k];
            if (a.apply_ls) x = (x - rmax) - rlog;
            const double val = exact_value(x, (allow[i * words + k] >> lane) & 1u, bsi);
            if (val >= tau) insert(i * V + lane + 32 * k, val);
          }
        }
      }
    }
  } else {
    scan([&](int c, double val) {
      if (val >= tau) insert(c, val);
    });
  }
  // ---- B2. nb rounds of warp arg-max over the list heads ---------------------------------------------------------
  // (value desc, flat index asc) on keys: warp max of the high word, of the low word among its holders, then the
  // smallest flat index among the holders of the full key
  int hd = 0;
  for (int j = 0; j < nb; ++j) {
    const unsigned long long key = hd < cnt ? f64_key(lv[hd]) : 0ull;      // 0 ranks below every real value
    const unsigned c_mine = hd < cnt ? (unsigned)lc[hd] : (unsigned)INT_MAX;
    const unsigned mh = __reduce_max_sync(0xffffffffu, (unsigned)(key >> 32));
    const bool top = (unsigned)(key >> 32) == mh;
    const unsigned ml = __reduce_max_sync(0xffffffffu, top ? (unsigned)key : 0u);
    const bool holder = top && (unsigned)key == ml;
    const unsigned c = __reduce_min_sync(0xffffffffu, holder ? c_mine : (unsigned)INT_MAX);
    if (hd < cnt && holder && c_mine == c) ++hd;                          // the owner pops its head
    if (lane == 0) {
      const unsigned long long wkey = ((unsigned long long)mh << 32) | ml;
      win_val[j] = wkey == 0ull ? -INFINITY : key_f64(wkey);
      win_idx[j] = (int)c;
    }
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
    const unsigned implicit = __ballot_sync(0xffffffffu, live && sj.node < 0);
#pragma unroll
    for (int j = 0; j < kWarpNb; ++j) {
      code[j] = INT_MAX;
      if ((implicit >> j) & 1u) {                                     // (warp-uniform)
        const int lo_j = __shfl_sync(0xffffffffu, sj.lo, j);
        const int n_j = __shfl_sync(0xffffffffu, nj, j);
        if (lane < n_j) code[j] = rb::trie_code(a.tv, (int64_t)lo_j + lane, t);
      }
    }
#pragma unroll
    for (int j = 0; j < kWarpNb; ++j) {
      if ((implicit >> j) & 1u) {
        const int v_j = __shfl_sync(0xffffffffu, vj, j);
        const int less = __popc(__ballot_sync(0xffffffffu, code[j] < v_j));
        const int leq = __popc(__ballot_sync(0xffffffffu, code[j] <= v_j));   // INT_MAX never counts (v_j < V)
        if (lane == j && leq != less) ns = TrieState{sj.lo + less, sj.lo + leq, -1, 0};
      }
    }
  }
  // every beam on a single leaf: the rest of the query's DocIDs is determined by the trie -> freeze it
  const bool frozen = a.allow_freeze &&
                      __all_sync(0xffffffffu, lane >= nb || (ns.hi - ns.lo == 1 && win_val[lane] > kPenalised));
  double* sc_out = frozen ? a.sc_fz : a.sc_new;
  TrieState* st_out = frozen ? a.st_fz : a.st_new;
  int32_t* hist_out = frozen ? a.hist_fz : a.hist_new;
  int32_t* anc_out = frozen ? a.anc_fz : a.anc_new;
  if (frozen && lane == 0) a.qstate[b] = t + 1;
  if (lane < nb) {
    const int r_new = b * nb + lane;
    sc_out[r_new] = win_val[lane];
    a.parent_out[r_new] = pj;
    a.token_out[r_new] = vj;
    st_out[r_new] = ns;
  }
  // token history and KV ancestry of the new beams: lane = position (L <= 32: one row per beam, no divisions, and the
  // parents' rows of ALL beams are requested before the first store so that their latencies overlap; longer DocIDs
  // walk the positions in chunks of 32)
  if (L <= 32) {
    int hv[kWarpNb], av[kWarpNb];
#pragma unroll
    for (int j = 0; j < kWarpNb; ++j) {
      if (j >= nb) break;
      const int pj_ = __shfl_sync(0xffffffffu, pj, j);
      const int src = b * nb + pj_;
      const bool ld = lane < t && lane < L;
      hv[j] = ld ? a.hist_old[src * L + lane] : 0;
      av[j] = ld ? a.anc_old[src * L + lane] : 0;
    }
#pragma unroll
    for (int j = 0; j < kWarpNb; ++j) {
      if (j >= nb) break;
      const int pj_ = __shfl_sync(0xffffffffu, pj, j), vj_ = __shfl_sync(0xffffffffu, vj, j);
      const int src = b * nb + pj_, dst = b * nb + j;
      if (lane < L) {
        const int p = lane;
        int anc = av[j];
        if (p == t) anc = (a.rpq == 1) ? b : src;
        else if (p == t + 1) anc = dst;
        hist_out[dst * L + p] = p == t ? vj_ : hv[j];
        anc_out[dst * L + p] = anc;
      }
    }
  } else {
    for (int j = 0; j < nb; ++j) {
      const int pj_ = __shfl_sync(0xffffffffu, pj, j), vj_ = __shfl_sync(0xffffffffu, vj, j);
      const int src = b * nb + pj_, dst = b * nb + j;
      for (int p = lane; p < L; p += 32) {
        hist_out[dst * L + p] = p < t ? a.hist_old[src * L + p] : (p == t ? vj_ : 0);
        int anc;
        if (p < t) anc = a.anc_old[src * L + p];
        else if (p == t) anc = (a.rpq == 1) ? b : src;
        else if (p == t + 1) anc = dst;
        else anc = 0;
        anc_out[dst * L + p] = anc;
      }
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
                                  int32_t* anc, int32_t* qstate, int32_t* qlist, int32_t* counts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < batch) { qstate[r] = 0; qlist[r] = r; }
  if (r == 0) { counts[0] = batch; counts[1] = 0; }
  if (r >= batch * nb) return;
  const int b = r / nb, i = r - b * nb;
  sc[r] = (i == 0) ? 0.0 : (double)(-1e9f);   // fp32 zeros with [:, 1:] = -1e9 (generation.py:418-420)
  st[r] = rb::trie_root(tv);
  for (int p = 0; p < L; ++p) { hist[r * L + p] = 0; anc[r * L + p] = (p == 0) ? b : 0; }
}

// one warp: stable compaction of the stepping list, newly frozen queries appended to the freeze-order list
__global__ void beam_compact_kernel(int nq_old, int32_t* qlist, const int32_t* qstate, int32_t* fz_list,
                                    int32_t* counts) {
  const int lane = threadIdx.x;
  int n_act = 0, n_fz = counts[1];
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < nq_old; base += 32) {
    const int idx = base + lane;
    const int q = idx < nq_old ? qlist[idx] : -1;
    const bool fz = q >= 0 && qstate[q] != 0;
    const bool act = q >= 0 && !fz;
    const unsigned ma = __ballot_sync(0xffffffffu, act), mf = __ballot_sync(0xffffffffu, fz);
    if (act) qlist[n_act + __popc(ma & lt)] = q;       // lands at or before idx: this chunk has been read already
    if (fz) fz_list[n_fz + __popc(mf & lt)] = q;
    n_act += __popc(ma);
    n_fz += __popc(mf);
  }
  if (lane == 0) { counts[0] = n_act; counts[1] = n_fz; }
}

__global__ void gather_rows_kernel(const int32_t* __restrict__ qlist, int nb, int d4, const float4* __restrict__ src,
                                   float4* __restrict__ dst) {
  const int m = blockIdx.x;
  const int r = qlist[m / nb] * nb + m % nb;
  for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[(int64_t)m * d4 + c] = src[(int64_t)r * d4 + c];
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
    // a NaN score (fp16x3 range overflow) still gets a rank of its own: every output slot is always written
    const double mine = hs[j] == hs[j] ? hs[j] : kNanRank;
    int rank = 0;
    for (int k = 0; k < nb; ++k) {
      const double other = hs[k] == hs[k] ? hs[k] : kNanRank;
      rank += (other > mine) || (other == mine && k > j);
    }
    if (rank >= keep) continue;
    const int64_t o = (int64_t)b * keep + rank;
    const int r = b * nb + j;
    seqs[o * (steps + 1)] = 0;   // decoder_start_token_id
    for (int p = 0; p < steps; ++p) seqs[o * (steps + 1) + 1 + p] = hist[r * L + p];
    scores[o] = (float)hs[j];
    if (leaf) { leaf[o * 2] = st[r].lo; leaf[o * 2 + 1] = st[r].hi; }
  }
}

// ---- forced tail ---------------------------------------------------------------------------------------------------
// one CTA per frozen row r' = fq * nb + i (freeze order)
__global__ void tail_prepare_kernel(TrieView tv, rb::TailLayout lay, int nb, int L, int d, int from_hist,
                                    const int32_t* __restrict__ fz_list, const int32_t* __restrict__ qstate,
                                    const TrieState* __restrict__ st, int32_t* __restrict__ hist,
                                    const float* const* __restrict__ in_tabs, const float* __restrict__ start_emb,
                                    float* __restrict__ x) {
  const int rp = blockIdx.x;
  const int fq = rp / nb, i = rp - fq * nb;
  const int b = fz_list[fq];
  const int r = b * nb + i;
  const int t0 = from_hist ? 0 : qstate[b];
  const int64_t leaf = st[r].lo;
  if (!from_hist) {
    for (int p = t0 + threadIdx.x; p < lay.P; p += blockDim.x) hist[r * L + p] = rb::trie_code(tv, leaf, p);
    __syncthreads();
  }
  if (x == nullptr) return;                      // (test hook: only the forced tokens are wanted)
  const int d4 = d >> 2;
  for (int p = t0; p < lay.P; ++p) {
    // input of position p = embedding of the token at p-1 (position 0: the learned start embedding)
    const float4* src = p == 0 ? reinterpret_cast<const float4*>(start_emb)
                               : reinterpret_cast<const float4*>(in_tabs[p - 1] + (int64_t)hist[r * L + p - 1] * d);
    float4* dst = reinterpret_cast<float4*>(x + ((int64_t)lay.off[p] + rp) * d);
    for (int c = threadIdx.x; c < d4; c += blockDim.x) dst[c] = src[c];
  }
}

// one warp per row of the pass: the forced token's (log-softmaxed) logit
__global__ void tail_pick_kernel(rb::TailLayout lay, int nb, int L, int V, int apply_ls,
                                 const int32_t* __restrict__ fz_list, const int32_t* __restrict__ hist,
                                 const float* __restrict__ logits, float* __restrict__ picked) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= lay.off[lay.P]) return;
  int p = 0;
  while (row >= lay.off[p + 1]) ++p;
  const int rp = row - lay.off[p];
  const int r = fz_list[rp / nb] * nb + rp % nb;
  const float* lr = logits + (int64_t)row * V;
  float x = lr[hist[r * L + p]];
  if (apply_ls) {
    float m = -INFINITY;
    for (int v = lane; v < V; v += 32) m = fmaxf(m, lr[v]);
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int v = lane; v < V; v += 32) sum += expf(lr[v] - m);
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    x = (x - m) - logf(sum);
  }
  if (lane == 0) picked[row] = x;
}

// One CTA per frozen query: replay its remaining steps. At each step the candidate of the beam in slot j is
// (score_j + x, flat index j * V + token) and the beams are re-ranked by (value desc, flat index asc), exactly what
// the step kernels do when every beam has a single valid child (all other candidates carry the -1e9 penalty).
__global__ void tail_finish_kernel(rb::TailLayout lay, int nb, int L, int V, int n2,
                                   const int32_t* __restrict__ fz_list,
                                   const int32_t* __restrict__ qstate, const float* __restrict__ picked,
                                   const double* __restrict__ sc_fz, const TrieState* __restrict__ st_fz,
                                   const int32_t* __restrict__ hist_fz, double* __restrict__ sc_out,
                                   TrieState* __restrict__ st_out, int32_t* __restrict__ hist_out) {
  // n2 = 0: rank counting (nb^2 compares per step, fine for a few dozen beams); n2 = nb rounded up to a power of two:
  // bitonic sort of (value desc, slot asc) per step - with 1000 beams 27 steps x 10^6 float64 compares on one SM took
  // 0.8 ms, the sorting network does the same re-ranking in 55 compare-exchange rounds per step
  extern __shared__ __align__(16) unsigned char tf_raw[];
  const int cap = n2 > 0 ? n2 : nb;
  double* s = reinterpret_cast<double*>(tf_raw);       // [nb] score of the beam in slot j
  double* val = s + cap;                               // [cap]
  int* perm = reinterpret_cast<int*>(val + cap);       // [nb] frozen beam index in slot j
  int* nperm = perm + cap;                             // [cap]: new order (rank counting) / slot of a sorted entry
  const int fq = blockIdx.x;
  const int b = fz_list[fq];
  const int t0 = qstate[b];
  for (int j = threadIdx.x; j < nb; j += blockDim.x) { s[j] = sc_fz[b * nb + j]; perm[j] = j; }
  __syncthreads();
  for (int p = t0; p < lay.P; ++p) {
    for (int j = threadIdx.x; j < cap; j += blockDim.x) {
      if (j < nb) {
        const double v = (double)picked[lay.off[p] + fq * nb + perm[j]] + s[j];
        val[j] = v == v ? v : kNanRank;
      } else {
        val[j] = -INFINITY;                              // padding of the sorting network: behind every real entry
      }
      if (n2 > 0) nperm[j] = j;
    }
    __syncthreads();
    if (n2 > 0) {
      // (value desc, slot asc); slot = position before this step = flat index order among equal values
      for (int k = 2; k <= n2; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
          for (int i = threadIdx.x; i < n2; i += blockDim.x) {
            const int q = i ^ jj;
            if (q > i) {
              const double vi = val[i], vq = val[q];
              const int si = nperm[i], sq = nperm[q];
              const bool i_first = vi > vq || (vi == vq && si < sq);
              const bool up = (i & k) == 0;
              if (up ? !i_first : i_first) {
                val[i] = vq; val[q] = vi;
                nperm[i] = sq; nperm[q] = si;
              }
            }
          }
          __syncthreads();
        }
      }
      int mine_perm = 0;
      for (int j = threadIdx.x; j < nb; j += blockDim.x) mine_perm = perm[nperm[j]];     // (nb <= blockDim.x on this path)
      __syncthreads();
      for (int j = threadIdx.x; j < nb; j += blockDim.x) { s[j] = val[j]; perm[j] = mine_perm; }
      __syncthreads();
      continue;
    }
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
      const double mine = val[j];
      // candidates ahead of slot j's under (value desc, flat index asc); flat index = slot * V + token with
      // token < V, so among equal values exactly the lower slots go first
      int rank = 0;
#pragma unroll 4
      for (int k = 0; k < j; ++k) rank += val[k] >= mine;
#pragma unroll 4
      for (int k = j + 1; k < nb; ++k) rank += val[k] > mine;
      s[rank] = mine;          // s[] is not read in this phase
      nperm[rank] = perm[j];
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nb; j += blockDim.x) perm[j] = nperm[j];
    __syncthreads();
  }
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    sc_out[b * nb + j] = s[j];
    st_out[b * nb + j] = st_fz[b * nb + perm[j]];
  }
  for (int e = threadIdx.x; e < nb * lay.P; e += blockDim.x) {
    const int j = e / lay.P, p = e - j * lay.P;
    hist_out[(b * nb + j) * L + p] = hist_fz[(b * nb + perm[j]) * L + p];
  }
}

__global__ void force_tokens_kernel(int batch, int nb, int L, int T, const int32_t* __restrict__ tokens,
                                    int32_t* __restrict__ hist, int32_t* __restrict__ fz_list,
                                    int32_t* __restrict__ qstate, int32_t* __restrict__ counts) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < batch) { fz_list[r] = r; qstate[r] = 0; }
  if (r == 0) { counts[0] = 0; counts[1] = batch; }
  if (r >= batch * nb) return;
  for (int p = 0; p < L; ++p) hist[r * L + p] = p < T ? tokens[r * T + p] : 0;
}

__global__ void forced_scores_kernel(rb::TailLayout lay, int R, int L, int V, const int32_t* __restrict__ hist,
                                     const float* __restrict__ logits, float* __restrict__ scores) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  double s = 0.0;
  for (int p = 0; p < lay.P; ++p) s += (double)logits[((int64_t)lay.off[p] + r) * V + hist[r * L + p]];
  scores[r] = (float)s;
}

size_t cta_smem_bytes(int nb, int words) {
  return (2 * nb + 32) * sizeof(double) + (32 + nb) * sizeof(int) + 2 * nb * sizeof(float) +
         (size_t)nb * words * sizeof(uint32_t) + (size_t)nb * sizeof(TrieState);
}
int select_pbuf(int nb) {
  int p = 1024;
  while (p < 2 * nb) p <<= 1;
  return p;
}
size_t select_smem_bytes(int nb, int words) {
  const int pbuf = select_pbuf(nb);
  return 2 * nb * sizeof(double) + (size_t)pbuf * 12 + nb * sizeof(int) + 2 * nb * sizeof(float) +
         kSelBins * sizeof(uint32_t) + (size_t)nb * words * sizeof(uint32_t) + (size_t)nb * sizeof(TrieState);
}

}  // namespace

namespace rb {

int beam_step(rb200_beam* bm, const rb200_trie* trie, const float* logits, int rows_per_query, int apply_log_softmax,
              const float* embed_table, float* next_x, int d_model, int allow_freeze, cudaStream_t stream) {
  RB_REQUIRE(bm && trie && logits, "null argument");
  RB_REQUIRE(bm->batch >= 1, "rb200_beam_reset has not been called");
  RB_REQUIRE(rows_per_query == 1 || rows_per_query == bm->nb, "rows_per_query must be 1 or num_beams=%d", bm->nb);
  RB_REQUIRE(bm->step < bm->L, "already took L=%d steps", bm->L);
  RB_REQUIRE(bm->step < trie->L, "step %d exceeds the trie depth %d", bm->step, trie->L);
  RB_REQUIRE((embed_table == nullptr) == (next_x == nullptr), "embed table and next_x must be given together");
  if (bm->n_active == 0) {           // every query is frozen: nothing steps any more
    bm->step += 1;
    return 0;
  }
  StepArgs a;
  a.tv = trie->device_view(bm->device);
  a.t = bm->step; a.nb = bm->nb; a.rpq = rows_per_query; a.apply_ls = apply_log_softmax; a.L = bm->L;
  a.d_model = d_model;
  a.nq = bm->n_active;
  a.qlist = bm->compacted ? bm->qlist : nullptr;
  a.allow_freeze = allow_freeze;
  a.logits = logits;
  const int o = bm->cur, n = bm->cur ^ 1;
  a.sc_old = bm->scores[o]; a.st_old = bm->state[o]; a.hist_old = bm->hist[o]; a.anc_old = bm->anc[o];
  a.sc_new = bm->scores[n]; a.st_new = bm->state[n]; a.hist_new = bm->hist[n]; a.anc_new = bm->anc[n];
  a.sc_fz = bm->fz_scores; a.st_fz = bm->fz_state; a.hist_fz = bm->fz_hist; a.anc_fz = bm->fz_anc;
  a.qstate = bm->qstate;
  a.parent_out = bm->parent; a.token_out = bm->token;
  a.embed_table = embed_table; a.next_x = next_x;
  const int nb = bm->nb;
  const int64_t total = (int64_t)nb * bm->V;
  const char* fe = getenv("RB200_BEAM");       // cta | select: force one formulation (parity tests); read per call
  const int force = !fe ? 0 : (strcmp(fe, "cta") == 0 ? 1 : (strcmp(fe, "select") == 0 ? 2 : 0));
  const size_t per_warp = warp_step_smem(a.tv.words, nb);
  const size_t smem = cta_smem_bytes(nb, a.tv.words);
  const bool cta_fits = total <= 1024 * kMaxPerThread && smem <= 200 * 1024;
  // (very wide codebooks would push the per-warp bitmaps past the default 48 KB of dynamic shared memory)
  if (force == 0 && nb <= kWarpNb && d_model % 4 == 0 && kWarpQ * per_warp <= 48 * 1024) {
    auto kern = a.tv.words == 8 && a.tv.V == 256 ? beam_step_warp_kernel<8>
                : (a.tv.words == 32 && a.tv.V == 1024 ? beam_step_warp_kernel<32> : beam_step_warp_kernel<0>);
    RB_CUDA(rb::launch_pdl(kern, dim3(rb::ceil_div(a.nq, kWarpQ)), dim3(kWarpQ * 32), kWarpQ * per_warp, stream, a));
  } else if (force != 2 && cta_fits && (force == 1 || total <= 256 * 32)) {
    // (beyond 8192 candidates per query the arg-max kernel rescans its 25+ candidates per thread after every
    // winner; the radix select is faster there: beam 100 x V 256 runs 249.5 ms per batch of 128 instead of 258.2)
    static bool attr_set = false;
    if (!attr_set) {
      RB_CUDA(cudaFuncSetAttribute(beam_step_kernel<256, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      RB_CUDA(cudaFuncSetAttribute(beam_step_kernel<256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      RB_CUDA(cudaFuncSetAttribute(beam_step_kernel<1024, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    if (total <= 256 * 16)
      RB_CUDA(rb::launch_pdl(beam_step_kernel<256, 16>, dim3(a.nq), dim3(256), smem, stream, a));
    else if (total <= 256 * 32)
      RB_CUDA(rb::launch_pdl(beam_step_kernel<256, 32>, dim3(a.nq), dim3(256), smem, stream, a));
    else
      RB_CUDA(rb::launch_pdl(beam_step_kernel<1024, 0>, dim3(a.nq), dim3(1024), smem, stream, a));
  } else {
    const size_t ssel = select_smem_bytes(nb, a.tv.words);
    RB_REQUIRE(nb <= kSelMaxNb && total < (1 << 24) && ssel <= kSelMaxSmem,
               "num_beams=%d x V=%d exceeds the beam kernels (needs %zu B of shared memory)", nb, bm->V, ssel);
    static bool attr_set = false;
    if (!attr_set) {
      RB_CUDA(cudaFuncSetAttribute(beam_step_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelMaxSmem));
      attr_set = true;
    }
    RB_CUDA(rb::launch_pdl(beam_step_select_kernel, dim3(a.nq), dim3(kSelThreads), ssel, stream, a, select_pbuf(nb)));
  }
  rb::launch_count()++;
  bm->cur = n;
  bm->step += 1;
  return 0;
}

int beam_compact(rb200_beam* bm, cudaStream_t s) {
  beam_compact_kernel<<<1, 32, 0, s>>>(bm->n_active, bm->qlist, bm->qstate, bm->fz_list, bm->counts);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

int launch_gather_rows(const rb200_beam* bm, const float* x_full, float* x, int d_model, cudaStream_t s) {
  const int M = bm->n_active * bm->nb;
  if (M == 0) return 0;
  gather_rows_kernel<<<M, 128, 0, s>>>(bm->qlist, bm->nb, d_model / 4, reinterpret_cast<const float4*>(x_full),
                                       reinterpret_cast<float4*>(x));
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

int launch_tail_prepare(rb200_beam* bm, const rb200_trie* trie, const TailLayout& lay, const float* const* in_tabs_dev,
                        const float* start_emb, float* x, int d_model, cudaStream_t s) {
  const int rows = bm->n_frozen * bm->nb;
  if (rows == 0) return 0;
  const bool from_hist = trie == nullptr;
  tail_prepare_kernel<<<rows, 128, 0, s>>>(from_hist ? TrieView{} : trie->device_view(bm->device), lay, bm->nb, bm->L, d_model,
                                           from_hist ? 1 : 0, bm->fz_list, bm->qstate, bm->fz_state, bm->fz_hist,
                                           in_tabs_dev, start_emb, x);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

int launch_tail_finish(rb200_beam* bm, const rb200_trie* trie, const TailLayout& lay, const float* logits,
                       int apply_log_softmax, cudaStream_t s) {
  (void)trie;
  const int nfz = bm->n_frozen, nb = bm->nb;
  if (nfz == 0) return 0;
  const int rows = lay.off[lay.P];
  float* picked = reinterpret_cast<float*>(bm->fz_anc);    // the ancestry of frozen queries is not needed any more
  RB_REQUIRE((int64_t)rows <= (int64_t)bm->max_batch * bm->max_nb * bm->L, "forced tail of %d rows exceeds the beam state", rows);
  tail_pick_kernel<<<ceil_div(rows, 4), 128, 0, s>>>(lay, nb, bm->L, bm->V, apply_log_softmax, bm->fz_list, bm->fz_hist,
                                                     logits, picked);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  const int c = bm->cur;
  int threads = ((nb + 31) / 32) * 32;
  threads = threads > 1024 ? 1024 : threads;
  // sorting network instead of rank counting for wide beams that fit one entry per thread (RB200_TAIL_SORT=0 | 1 forces)
  int n2 = 0;
  {
    const char* e = getenv("RB200_TAIL_SORT");
    const bool want = e ? e[0] == '1' : nb > 64;
    if (want && nb <= 1024) {
      n2 = 2;
      while (n2 < nb) n2 <<= 1;
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    RB_CUDA(cudaFuncSetAttribute(tail_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelMaxNb * 24));
    attr_set = true;
  }
  const int cap = n2 > 0 ? n2 : nb;
  tail_finish_kernel<<<nfz, threads, (size_t)cap * 24, s>>>(lay, nb, bm->L, bm->V, n2, bm->fz_list, bm->qstate, picked,
                                                           bm->fz_scores, bm->fz_state, bm->fz_hist, bm->scores[c],
                                                           bm->state[c], bm->hist[c]);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

int beam_force_tokens(rb200_beam* bm, int batch, int nb, int T, const int32_t* tokens_dev, cudaStream_t s) {
  RB_REQUIRE(batch >= 1 && batch <= bm->max_batch && nb >= 1 && nb <= bm->max_nb, "batch %d x rows %d outside the beam state %d x %d",
             batch, nb, bm->max_batch, bm->max_nb);
  RB_REQUIRE(T >= 1 && T <= bm->L, "T=%d outside [1, %d]", T, bm->L);
  bm->batch = batch; bm->nb = nb; bm->step = 0; bm->cur = 0;
  bm->n_active = 0; bm->n_frozen = batch; bm->compacted = false;
  const int R = batch * nb;
  force_tokens_kernel<<<ceil_div(R, 128), 128, 0, s>>>(batch, nb, bm->L, T, tokens_dev, bm->fz_hist, bm->fz_list,
                                                       bm->qstate, bm->counts);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

int launch_forced_scores(const rb200_beam* bm, const TailLayout& lay, const float* logits, float* scores,
                         cudaStream_t s) {
  const int R = bm->batch * bm->nb;
  forced_scores_kernel<<<ceil_div(R, 128), 128, 0, s>>>(lay, R, bm->L, bm->V, bm->fz_hist, logits, scores);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

}  // namespace rb

extern "C" {

int rb200_beam_create(int device, int max_batch, int num_beams, int L, int V, rb200_beam** out) {
  RB_REQUIRE(out, "null argument");
  RB_REQUIRE(max_batch >= 1 && num_beams >= 1 && L >= 1 && V >= 1, "need max_batch, num_beams, L, V >= 1");
  const int words = (V + 31) / 32;
  const bool fits = ((int64_t)num_beams * V <= 1024 * kMaxPerThread && cta_smem_bytes(num_beams, words) <= 200 * 1024) ||
                    (num_beams <= kSelMaxNb && (int64_t)num_beams * V < (1 << 24) &&
                     select_smem_bytes(num_beams, words) <= kSelMaxSmem);
  RB_REQUIRE(fits, "num_beams=%d x V=%d exceeds the beam kernels' shared-memory budget", num_beams, V);
  rb200_beam* bm = new (std::nothrow) rb200_beam();
  if (!bm) return rb::fail(RB200_ERR_NOMEM, "out of memory");
  bm->device = device; bm->max_batch = max_batch; bm->max_nb = num_beams; bm->nb = num_beams; bm->L = L; bm->V = V;
  int prev = 0;
  cudaError_t err = cudaGetDevice(&prev);
  if (err == cudaSuccess) err = cudaSetDevice(device);
  const size_t R = (size_t)max_batch * num_beams;
  auto alloc = [&](void** p, size_t bytes) {
    if (err == cudaSuccess) err = cudaMalloc(p, bytes);
  };
  for (int h = 0; h < 2; ++h) {
    alloc((void**)&bm->scores[h], R * sizeof(double));
    alloc((void**)&bm->state[h], R * sizeof(TrieState));
    alloc((void**)&bm->hist[h], R * L * sizeof(int32_t));
    alloc((void**)&bm->anc[h], R * L * sizeof(int32_t));
  }
  alloc((void**)&bm->parent, R * sizeof(int32_t));
  alloc((void**)&bm->token, R * sizeof(int32_t));
  alloc((void**)&bm->qstate, (size_t)max_batch * sizeof(int32_t));
  alloc((void**)&bm->qlist, (size_t)max_batch * sizeof(int32_t));
  alloc((void**)&bm->fz_list, (size_t)max_batch * sizeof(int32_t));
  alloc((void**)&bm->counts, 2 * sizeof(int32_t));
  alloc((void**)&bm->fz_scores, R * sizeof(double));
  alloc((void**)&bm->fz_state, R * sizeof(TrieState));
  alloc((void**)&bm->fz_hist, R * L * sizeof(int32_t));
  alloc((void**)&bm->fz_anc, R * L * sizeof(int32_t));
  if (err != cudaSuccess) {          // give back whatever was allocated (cudaFree(nullptr) is a no-op)
    rb200_beam_free(bm);
    cudaSetDevice(prev);
    return rb::fail(RB200_ERR_CUDA, "rb200_beam_create: %s", cudaGetErrorString(err));
  }
  cudaSetDevice(prev);
  *out = bm;
  return 0;
}

int rb200_beam_free(rb200_beam* bm) {
  if (!bm) return 0;
  for (int h = 0; h < 2; ++h) {
    cudaFree(bm->scores[h]); cudaFree(bm->state[h]); cudaFree(bm->hist[h]); cudaFree(bm->anc[h]);
  }
  cudaFree(bm->parent); cudaFree(bm->token);
  cudaFree(bm->qstate); cudaFree(bm->qlist); cudaFree(bm->fz_list); cudaFree(bm->counts);
  cudaFree(bm->fz_scores); cudaFree(bm->fz_state); cudaFree(bm->fz_hist); cudaFree(bm->fz_anc);
  delete bm;
  return 0;
}

int rb200_beam_reset_beams(rb200_beam* bm, const rb200_trie* trie, int batch, int num_beams, void* stream) {
  RB_REQUIRE(bm && trie, "null argument");
  RB_REQUIRE(batch >= 1 && batch <= bm->max_batch, "batch %d outside [1, %d]", batch, bm->max_batch);
  RB_REQUIRE(num_beams >= 1 && num_beams <= bm->max_nb, "num_beams %d outside [1, %d]", num_beams, bm->max_nb);
  RB_REQUIRE(trie->V == bm->V, "trie V=%d but beam state was created for V=%d", trie->V, bm->V);
  if (trie->tables_on(bm->device) == nullptr)
    return rb::fail(RB200_ERR_STATE, "the trie has not been uploaded to device %d (beam state): call rb200_trie_upload",
                    bm->device);
  bm->batch = batch; bm->nb = num_beams; bm->step = 0; bm->cur = 0;
  bm->n_active = batch; bm->n_frozen = 0; bm->compacted = false;
  const int R = batch * bm->nb;
  beam_reset_kernel<<<rb::ceil_div(R, 128), 128, 0, (cudaStream_t)stream>>>(
      trie->device_view(bm->device), bm->nb, bm->L, batch, bm->scores[0], bm->state[0], bm->hist[0], bm->anc[0], bm->qstate,
      bm->qlist, bm->counts);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

int rb200_beam_reset(rb200_beam* bm, const rb200_trie* trie, int batch, void* stream) {
  RB_REQUIRE(bm, "null argument");
  return rb200_beam_reset_beams(bm, trie, batch, bm->max_nb, stream);
}

int rb200_beam_step(rb200_beam* bm, const rb200_trie* trie, const float* logits, int rows_per_query,
                    int apply_log_softmax, const float* embed_table, float* next_x, int d_model, void* stream) {
  return rb::beam_step(bm, trie, logits, rows_per_query, apply_log_softmax, embed_table, next_x, d_model, 0,
                       (cudaStream_t)stream);
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

// Test hook for the forced tail's score replay: the beam state must sit at step t with every beam of every query on
// a single leaf (as after rb200_beam_step calls that forced them). tail_logits_dev: fp32 [T, batch*nb, V], block j =
// the logits the decoder would produce at step t + j for the beams IN THEIR ORDER AT STEP t (a lineage's logits do
// not depend on the slot it occupies). Leaves the state exactly where T more rb200_beam_step calls would.
int rb200_beam_forced_tail(rb200_beam* bm, const rb200_trie* trie, int T, const float* tail_logits_dev,
                           int apply_log_softmax, void* stream) {
  RB_REQUIRE(bm && trie && tail_logits_dev, "null argument");
  RB_REQUIRE(bm->batch >= 1 && bm->step >= 1, "no step has been taken");
  RB_REQUIRE(T >= 1 && bm->step + T <= bm->L && bm->step + T <= trie->L && bm->step + T <= RB_TAIL_MAX_L,
             "T=%d steps after step %d exceed min(L, %d)", T, bm->step, RB_TAIL_MAX_L);
  cudaStream_t s = (cudaStream_t)stream;
  const int R = bm->batch * bm->nb, c = bm->cur, t0 = bm->step;
  // freeze every query at the current step: fz_* <- current state, freeze order = query order
  RB_CUDA(cudaMemcpyAsync(bm->fz_scores, bm->scores[c], (size_t)R * sizeof(double), cudaMemcpyDeviceToDevice, s));
  RB_CUDA(cudaMemcpyAsync(bm->fz_state, bm->state[c], (size_t)R * sizeof(TrieState), cudaMemcpyDeviceToDevice, s));
  RB_CUDA(cudaMemcpyAsync(bm->fz_hist, bm->hist[c], (size_t)R * bm->L * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  std::vector<int32_t> ident(bm->batch), st(bm->batch, t0);
  for (int i = 0; i < bm->batch; ++i) ident[i] = i;
  RB_CUDA(cudaMemcpyAsync(bm->fz_list, ident.data(), ident.size() * 4, cudaMemcpyHostToDevice, s));
  RB_CUDA(cudaMemcpyAsync(bm->qstate, st.data(), st.size() * 4, cudaMemcpyHostToDevice, s));
  RB_CUDA(cudaStreamSynchronize(s));
  bm->n_frozen = bm->batch; bm->n_active = 0;
  rb::TailLayout lay;
  lay.P = t0 + T;
  for (int p = 0; p <= RB_TAIL_MAX_L; ++p) lay.off[p] = p <= t0 ? 0 : (p - t0) * R;
  // the forced tokens of the remaining positions -> fz_hist (x = nullptr: no decoder inputs are wanted here)
  RB_TRY(rb::launch_tail_prepare(bm, trie, lay, nullptr, nullptr, nullptr, 4, s));
  RB_TRY(rb::launch_tail_finish(bm, trie, lay, tail_logits_dev, apply_log_softmax, s));
  bm->step += T;
  bm->n_active = bm->batch; bm->n_frozen = 0;      // the state is whole again in the current half
  RB_CUDA(cudaMemsetAsync(bm->qstate, 0, (size_t)bm->batch * 4, s));
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
