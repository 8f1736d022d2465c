// Flattened DocID trie over the lexicographically sorted unique code rows.
//
// A trie node at depth t is a contiguous range [lo, hi) of sorted unique codes sharing a prefix of
// length t. Three regimes, all reachable from the same (lo, hi, node) beam state:
//   * hi - lo == 0        dead: the prefix is not in the trie (the reference's all-zero mask row,
//                         generation.py:656-661,675);
//   * node >= 0           explicit node (ranges holding > RB_TRIE_SMALL codes): V-bit child bitmap +
//                         child range table; child k is found by popcount rank;
//   * 1..RB_TRIE_SMALL    implicit: the children are read straight from column t of the <= 32 code rows
//                         (covers the long unary tails: 8.8M x 32 codes would otherwise need ~2.6e8 nodes).
#pragma once
#include <cstdint>
#include <vector>

#ifdef __CUDACC__
#define RB_HD __host__ __device__ __forceinline__
#else
#define RB_HD inline
#endif

#define RB_TRIE_SMALL 32

namespace rb {

struct TrieView {
  const uint8_t* codes8 = nullptr;      // [U, L] when code_bytes == 1
  const uint16_t* codes16 = nullptr;    // [U, L] when code_bytes == 2
  const uint32_t* node_bitmap = nullptr;  // [n_nodes, words]
  const int32_t* node_child_ptr = nullptr;  // [n_nodes] -> first slot in child_lo / child_node
  const int32_t* child_lo = nullptr;    // per node: count+1 entries (last = node hi)
  const int32_t* child_node = nullptr;  // explicit node id of the child or -1
  int32_t L = 0, V = 0, words = 0, U = 0, root_node = -1;
};

struct TrieState {
  int32_t lo, hi, node, pad;
};

RB_HD int trie_code(const TrieView& tv, int64_t row, int t) {
  return tv.codes8 ? (int)tv.codes8[row * tv.L + t] : (int)tv.codes16[row * tv.L + t];
}

RB_HD TrieState trie_root(const TrieView& tv) { return TrieState{0, tv.U, tv.root_node, 0}; }
RB_HD TrieState trie_dead() { return TrieState{0, 0, -1, 0}; }

RB_HD int rb_popc(uint32_t x) {
#ifdef __CUDA_ARCH__
  return __popc(x);
#else
  return __builtin_popcount(x);
#endif
}

// Allowed next tokens of state s at depth t as a V-bit bitmap (tv.words 32-bit words).
RB_HD void trie_allowed(const TrieView& tv, const TrieState& s, int t, uint32_t* bitmap) {
  for (int w = 0; w < tv.words; ++w) bitmap[w] = 0u;
  const int n = s.hi - s.lo;
  if (n <= 0 || t >= tv.L) return;
  if (s.node >= 0) {
    const uint32_t* src = tv.node_bitmap + (int64_t)s.node * tv.words;
    for (int w = 0; w < tv.words; ++w) bitmap[w] = src[w];
    return;
  }
  for (int j = 0; j < n; ++j) {
    const int v = trie_code(tv, (int64_t)s.lo + j, t);
    bitmap[v >> 5] |= 1u << (v & 31);
  }
}

// State after taking token v at depth t; dead if v is not a child.
RB_HD TrieState trie_child(const TrieView& tv, const TrieState& s, int t, int v) {
  const int n = s.hi - s.lo;
  if (n <= 0 || t >= tv.L || v < 0 || v >= tv.V) return trie_dead();
  if (s.node >= 0) {
    const uint32_t* bm = tv.node_bitmap + (int64_t)s.node * tv.words;
    const int w = v >> 5;
    const uint32_t bit = 1u << (v & 31);
    if (!(bm[w] & bit)) return trie_dead();
    int k = rb_popc(bm[w] & (bit - 1u));
    for (int i = 0; i < w; ++i) k += rb_popc(bm[i]);
    const int c = tv.node_child_ptr[s.node] + k;
    return TrieState{tv.child_lo[c], tv.child_lo[c + 1], tv.child_node[c], 0};
  }
  int less = 0, leq = 0;
  for (int j = 0; j < n; ++j) {
    const int c = trie_code(tv, (int64_t)s.lo + j, t);
    less += (c < v);
    leq += (c <= v);
  }
  if (leq == less) return trie_dead();
  return TrieState{s.lo + less, s.lo + leq, -1, 0};
}

}  // namespace rb

// The opaque handle of the C ABI.
struct rb200_trie {
  int32_t L = 0, V = 0, code_bytes = 1, words = 0;
  int64_t n_docs = 0, U = 0;
  std::vector<uint8_t> codes;          // U * L * code_bytes, sorted unique rows
  std::vector<uint32_t> node_bitmap;
  std::vector<int32_t> node_child_ptr;
  std::vector<int32_t> child_lo;
  std::vector<int32_t> child_node;
  std::vector<int64_t> leaf_ptr;       // [U+1]
  std::vector<int64_t> leaf_docs;      // [n_docs] input row indices grouped by leaf, input order inside
  std::vector<int64_t> level_counts;   // [L]
  int32_t root_node = -1;
  // device copies: one set of tables per device the trie was uploaded to (a single-process multi-GPU caller uploads
  // the same handle to every device; `device` is the most recent one, reported by rb200_trie_get_info)
  struct DevTables {
    int device = -1;
    void* codes = nullptr;
    uint32_t* node_bitmap = nullptr;
    int32_t* node_child_ptr = nullptr;
    int32_t* child_lo = nullptr;
    int32_t* child_node = nullptr;
    int32_t* leaf_ptr = nullptr;       // [U+1] (rb200_trie_leaf_expand; uploaded on first use)
    int32_t* leaf_docs = nullptr;      // [n_docs]
  };
  std::vector<DevTables> dev;
  int device = -1;

  const DevTables* tables_on(int d) const {
    for (const auto& t : dev)
      if (t.device == d) return &t;
    return nullptr;
  }
  DevTables* tables_on(int d) {
    for (auto& t : dev)
      if (t.device == d) return &t;
    return nullptr;
  }
  rb::TrieView host_view() const;
  rb::TrieView device_view(int d) const;      // tables of device d (must have been uploaded there)
  int64_t table_bytes() const;
};
