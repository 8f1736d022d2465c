// DocID trie: host builder, binary cache, host walk and the device mask kernel.
// Replaces reference evaluate.py:404-446 (dict building) and generation.py:604-677 (mask tables + call).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <unistd.h>

#include "rb_common.h"
#include "trie.h"

using rb::TrieState;
using rb::TrieView;

rb::TrieView rb200_trie::host_view() const {
  TrieView v;
  if (code_bytes == 1) v.codes8 = codes.data();
  else v.codes16 = reinterpret_cast<const uint16_t*>(codes.data());
  v.node_bitmap = node_bitmap.data();
  v.node_child_ptr = node_child_ptr.data();
  v.child_lo = child_lo.data();
  v.child_node = child_node.data();
  v.L = L; v.V = V; v.words = words; v.U = (int32_t)U; v.root_node = root_node;
  return v;
}

rb::TrieView rb200_trie::device_view(int d) const {
  TrieView v;
  const DevTables* t = tables_on(d);
  if (t == nullptr) return v;
  if (code_bytes == 1) v.codes8 = static_cast<const uint8_t*>(t->codes);
  else v.codes16 = static_cast<const uint16_t*>(t->codes);
  v.node_bitmap = t->node_bitmap;
  v.node_child_ptr = t->node_child_ptr;
  v.child_lo = t->child_lo;
  v.child_node = t->child_node;
  v.L = L; v.V = V; v.words = words; v.U = (int32_t)U; v.root_node = root_node;
  return v;
}

int64_t rb200_trie::table_bytes() const {
  return (int64_t)codes.size() + 4 * (int64_t)(node_bitmap.size() + node_child_ptr.size() + child_lo.size() +
                                               child_node.size());
}

namespace {

template <typename T>
int build_impl(const T* codes, int64_t N, int L, int V, int n_threads, rb200_trie* tr) {
  for (int64_t i = 0; i < N * L; ++i)
    if ((int)codes[i] >= V) return rb::fail(RB200_ERR_INVALID, "code %d at flat index %lld is >= V=%d",
                                            (int)codes[i], (long long)i, V);
  // 1. stable lexicographic argsort, parallel over first-code buckets.
  std::vector<int64_t> order(N);
  std::vector<int64_t> bucket_start(V + 1, 0);
  for (int64_t i = 0; i < N; ++i) bucket_start[codes[i * L] + 1]++;
  for (int v = 0; v < V; ++v) bucket_start[v + 1] += bucket_start[v];
  {
    std::vector<int64_t> cur(bucket_start.begin(), bucket_start.end() - 1);
    for (int64_t i = 0; i < N; ++i) order[cur[codes[i * L]]++] = i;   // stable within bucket
  }
  auto less = [&](int64_t a, int64_t b) {
    const T* pa = codes + a * L;
    const T* pb = codes + b * L;
    if (sizeof(T) == 1) {
      int c = memcmp(pa, pb, L);
      if (c != 0) return c < 0;
    } else {
      for (int k = 0; k < L; ++k)
        if (pa[k] != pb[k]) return pa[k] < pb[k];
    }
    return a < b;
  };
  int nt = std::max(1, std::min(n_threads <= 0 ? (int)std::thread::hardware_concurrency() : n_threads, 64));
  {
    std::vector<std::thread> th;
    std::atomic<int> next{0};
    for (int w = 0; w < nt; ++w)
      th.emplace_back([&]() {
        for (;;) {
          int v = next.fetch_add(1);
          if (v >= V) break;
          std::sort(order.begin() + bucket_start[v], order.begin() + bucket_start[v + 1], less);
        }
      });
    for (auto& t : th) t.join();
  }
  // 2. unique rows, leaf CSR, level counts.
  auto same = [&](int64_t a, int64_t b) { return memcmp(codes + a * L, codes + b * L, sizeof(T) * L) == 0; };
  tr->leaf_docs.assign(order.begin(), order.end());
  tr->leaf_ptr.clear();
  tr->leaf_ptr.push_back(0);
  std::vector<int64_t> firsts;
  for (int64_t i = 0; i < N; ++i) {
    if (i == 0 || !same(order[i - 1], order[i])) {
      if (i != 0) tr->leaf_ptr.push_back(i);
      firsts.push_back(order[i]);
    }
  }
  tr->leaf_ptr.push_back(N);
  const int64_t U = (int64_t)firsts.size();
  if (U > 0x7fffffff) return rb::fail(RB200_ERR_INVALID, "too many unique codes (%lld)", (long long)U);
  tr->U = U;
  tr->codes.resize((size_t)U * L * sizeof(T));
  T* uc = reinterpret_cast<T*>(tr->codes.data());
  for (int64_t u = 0; u < U; ++u) memcpy(uc + u * L, codes + firsts[u] * L, sizeof(T) * L);
  // a row whose longest common prefix with its predecessor is lcp starts a new prefix at every length > lcp
  std::vector<int64_t> lcp_hist(L + 1, 0);
  for (int64_t u = 1; u < U; ++u) {
    int lcp = 0;
    while (lcp < L && uc[(u - 1) * L + lcp] == uc[u * L + lcp]) ++lcp;
    lcp_hist[lcp]++;
  }
  tr->level_counts.assign(L, 1);
  int64_t acc = 0;
  for (int i = 1; i < L; ++i) {
    acc += lcp_hist[i - 1];
    tr->level_counts[i] = 1 + acc;
  }
  // 3. explicit nodes (BFS) for ranges holding more than RB_TRIE_SMALL codes.
  tr->words = (V + 31) / 32;
  struct Item { int32_t lo, hi, depth; };
  std::vector<Item> queue;
  tr->root_node = -1;
  if (U > RB_TRIE_SMALL && L > 0) {
    queue.push_back({0, (int32_t)U, 0});
    tr->root_node = 0;
  }
  for (size_t q = 0; q < queue.size(); ++q) {
    const Item it = queue[q];
    const size_t bm_off = tr->node_bitmap.size();
    tr->node_bitmap.resize(bm_off + tr->words, 0u);
    tr->node_child_ptr.push_back((int32_t)tr->child_lo.size());
    int32_t start = it.lo;
    while (start < it.hi) {
      const int v = uc[(int64_t)start * L + it.depth];
      // upper bound of v in column `depth` over [start, hi): the column is non-decreasing in the range
      int32_t a = start + 1, b = it.hi;
      while (a < b) {
        int32_t m = a + (b - a) / 2;
        if ((int)uc[(int64_t)m * L + it.depth] <= v) a = m + 1; else b = m;
      }
      const int32_t end = a;
      tr->node_bitmap[bm_off + (v >> 5)] |= 1u << (v & 31);
      tr->child_lo.push_back(start);
      if (end - start > RB_TRIE_SMALL && it.depth + 1 < L) {
        tr->child_node.push_back((int32_t)queue.size());
        queue.push_back({start, end, it.depth + 1});
      } else {
        tr->child_node.push_back(-1);
      }
      start = end;
    }
    tr->child_lo.push_back(it.hi);      // sentinel: hi of the last child
    tr->child_node.push_back(-1);
  }
  if (tr->node_bitmap.empty()) {         // keep device pointers non-null
    tr->node_bitmap.assign(tr->words, 0u);
    tr->node_child_ptr.assign(1, 0);
    tr->child_lo.assign(2, 0);
    tr->child_node.assign(2, -1);
  }
  return 0;
}

__global__ void trie_mask_kernel(TrieView tv, const int64_t* __restrict__ ids, int64_t R, int T,
                                 double* __restrict__ mask) {
  // one warp per row: lane 0 walks the prefix, all lanes write the V mask values coalesced.
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= R) return;
  extern __shared__ uint32_t smem_bm[];
  uint32_t* bm = smem_bm + (threadIdx.x >> 5) * tv.words;
  if (lane == 0) {
    TrieState s = rb::trie_root(tv);
    for (int t = 1; t < T; ++t) s = rb::trie_child(tv, s, t - 1, (int)ids[row * T + t]);
    rb::trie_allowed(tv, s, T - 1, bm);
  }
  __syncwarp();
  for (int v = lane; v < tv.V; v += 32) mask[row * tv.V + v] = (bm[v >> 5] >> (v & 31)) & 1u ? 1.0 : 0.0;
}

// One warp per output row: the input rows (documents) under the row's leaf range [lo, hi), ascending = the order of
// docid_to_smtid.json (evaluate.py:439-446 appends docids in json iteration order). leaf_docs is grouped by leaf with
// input order inside a leaf, so a single leaf is copied as is; a range of several leaves (prefix search) is rank-sorted.
__global__ void leaf_expand_kernel(const int32_t* __restrict__ leaf_ptr, const int32_t* __restrict__ leaf_docs,
                                   const int32_t* __restrict__ ranges, int64_t n, int k, int64_t* __restrict__ docs,
                                   int32_t* __restrict__ counts) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  const int lo = ranges[row * 2], hi = ranges[row * 2 + 1];
  const int beg = hi > lo ? leaf_ptr[lo] : 0, end = hi > lo ? leaf_ptr[hi] : 0;
  const int cnt = end - beg;
  if (lane == 0) counts[row] = cnt;
  int64_t* out = docs + row * k;
  if (cnt > k) {                                   // does not fit: the caller expands this row on the host
    for (int e = lane; e < k; e += 32) out[e] = -1;
    return;
  }
  for (int e = lane; e < k; e += 32) {
    if (e >= cnt) { out[e] = -1; continue; }
    const int v = leaf_docs[beg + e];
    int rank = e;
    if (hi - lo > 1) {
      rank = 0;
      for (int j = 0; j < cnt; ++j) rank += leaf_docs[beg + j] < v;      // row ids are distinct
    }
    out[rank] = v;
  }
}

static const char kMagic[8] = {'R', 'B', '2', 'T', 'R', 'I', 'E', '2'};

template <typename T>
static bool wr(FILE* f, const std::vector<T>& v) {
  int64_t n = (int64_t)v.size();
  return fwrite(&n, 8, 1, f) == 1 && (n == 0 || fwrite(v.data(), sizeof(T), n, f) == (size_t)n);
}
template <typename T>
static bool rd(FILE* f, std::vector<T>& v) {
  int64_t n = 0;
  if (fread(&n, 8, 1, f) != 1 || n < 0) return false;
  v.resize(n);
  return n == 0 || fread(v.data(), sizeof(T), n, f) == (size_t)n;
}

}  // namespace

extern "C" {

int rb200_trie_build(const void* codes_host, int code_bytes, int64_t n_docs, int L, int V, int n_threads,
                     rb200_trie** out) {
  RB_REQUIRE(codes_host && out, "null argument");
  RB_REQUIRE(code_bytes == 1 || code_bytes == 2, "code_bytes must be 1 or 2, got %d", code_bytes);
  RB_REQUIRE(n_docs >= 1 && L >= 1 && V >= 1, "need n_docs, L, V >= 1");
  RB_REQUIRE(V <= (code_bytes == 1 ? 256 : 65536), "V=%d does not fit %d-byte codes", V, code_bytes);
  rb200_trie* tr = new (std::nothrow) rb200_trie();
  if (!tr) return rb::fail(RB200_ERR_NOMEM, "out of memory");
  tr->L = L; tr->V = V; tr->code_bytes = code_bytes; tr->n_docs = n_docs;
  int st = code_bytes == 1 ? build_impl(static_cast<const uint8_t*>(codes_host), n_docs, L, V, n_threads, tr)
                           : build_impl(static_cast<const uint16_t*>(codes_host), n_docs, L, V, n_threads, tr);
  if (st != 0) { delete tr; return st; }
  *out = tr;
  return 0;
}

int rb200_trie_free(rb200_trie* tr) {
  if (!tr) return 0;
  int prev = 0;
  const bool have_dev = !tr->dev.empty() && cudaGetDevice(&prev) == cudaSuccess;
  for (auto& t : tr->dev) {
    cudaSetDevice(t.device);
    cudaFree(t.codes); cudaFree(t.node_bitmap); cudaFree(t.node_child_ptr);
    cudaFree(t.child_lo); cudaFree(t.child_node);
    cudaFree(t.leaf_ptr); cudaFree(t.leaf_docs);
  }
  if (have_dev) cudaSetDevice(prev);
  delete tr;
  return 0;
}

int rb200_trie_get_info(const rb200_trie* tr, rb200_trie_info* info) {
  RB_REQUIRE(tr && info, "null argument");
  info->n_docs = tr->n_docs; info->n_unique = tr->U;
  info->n_nodes = (int64_t)tr->node_child_ptr.size();
  info->n_children = (int64_t)tr->child_lo.size();
  info->bytes = tr->table_bytes();
  info->L = tr->L; info->V = tr->V; info->code_bytes = tr->code_bytes; info->on_device = tr->device;
  return 0;
}

int rb200_trie_level_counts(const rb200_trie* tr, int64_t* counts) {
  RB_REQUIRE(tr && counts, "null argument");
  for (int i = 0; i < tr->L; ++i) counts[i] = tr->level_counts[i];
  return 0;
}

int rb200_trie_save_tagged(const rb200_trie* tr, const char* path, uint64_t source_tag) {
  RB_REQUIRE(tr && path, "null argument");
  // written next to the target and renamed into place: a reader never sees a half-written cache
  const std::string tmp = std::string(path) + ".tmp." + std::to_string((long long)getpid());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return rb::fail(RB200_ERR_IO, "cannot open %s for writing", tmp.c_str());
  int64_t hdr[7] = {tr->L, tr->V, tr->code_bytes, tr->n_docs, tr->U, tr->root_node, (int64_t)source_tag};
  bool ok = fwrite(kMagic, 8, 1, f) == 1 && fwrite(hdr, 8, 7, f) == 7 && wr(f, tr->codes) &&
            wr(f, tr->node_bitmap) && wr(f, tr->node_child_ptr) && wr(f, tr->child_lo) && wr(f, tr->child_node) &&
            wr(f, tr->leaf_ptr) && wr(f, tr->leaf_docs) && wr(f, tr->level_counts);
  ok = (fclose(f) == 0) && ok;
  if (ok) ok = rename(tmp.c_str(), path) == 0;
  if (!ok) {
    remove(tmp.c_str());
    return rb::fail(RB200_ERR_IO, "short write to %s", path);
  }
  return 0;
}

int rb200_trie_save(const rb200_trie* tr, const char* path) { return rb200_trie_save_tagged(tr, path, 0); }

int rb200_trie_load_tagged(const char* path, uint64_t* source_tag, rb200_trie** out) {
  RB_REQUIRE(path && out, "null argument");
  FILE* f = fopen(path, "rb");
  if (!f) return rb::fail(RB200_ERR_IO, "cannot open %s", path);
  char magic[8];
  int64_t hdr[7];
  rb200_trie* tr = new rb200_trie();
  bool ok = fread(magic, 8, 1, f) == 1 && memcmp(magic, kMagic, 8) == 0 && fread(hdr, 8, 7, f) == 7;
  if (ok) {
    tr->L = (int)hdr[0]; tr->V = (int)hdr[1]; tr->code_bytes = (int)hdr[2]; tr->n_docs = hdr[3]; tr->U = hdr[4];
    tr->root_node = (int)hdr[5];
    if (source_tag) *source_tag = (uint64_t)hdr[6];
    tr->words = (tr->V + 31) / 32;
    ok = rd(f, tr->codes) && rd(f, tr->node_bitmap) && rd(f, tr->node_child_ptr) && rd(f, tr->child_lo) &&
         rd(f, tr->child_node) && rd(f, tr->leaf_ptr) && rd(f, tr->leaf_docs) && rd(f, tr->level_counts);
    ok = ok && (int64_t)tr->codes.size() == tr->U * tr->L * tr->code_bytes &&
         (int64_t)tr->leaf_ptr.size() == tr->U + 1 && (int64_t)tr->leaf_docs.size() == tr->n_docs;
  }
  fclose(f);
  if (!ok) { delete tr; return rb::fail(RB200_ERR_IO, "%s is not a riporb200 trie file (or is truncated)", path); }
  *out = tr;
  return 0;
}

int rb200_trie_load(const char* path, rb200_trie** out) { return rb200_trie_load_tagged(path, nullptr, out); }

int rb200_trie_mask_host(const rb200_trie* tr, const int64_t* ids, int64_t R, int T, double* mask) {
  RB_REQUIRE(tr && ids && mask, "null argument");
  RB_REQUIRE(T >= 1 && T <= tr->L, "prefix length T=%d outside [1, L=%d]", T, tr->L);
  const TrieView tv = tr->host_view();
  std::vector<uint32_t> bm(tv.words);
  for (int64_t r = 0; r < R; ++r) {
    TrieState s = rb::trie_root(tv);
    for (int t = 1; t < T; ++t) s = rb::trie_child(tv, s, t - 1, (int)ids[r * T + t]);
    rb::trie_allowed(tv, s, T - 1, bm.data());
    for (int v = 0; v < tv.V; ++v) mask[r * tv.V + v] = (bm[v >> 5] >> (v & 31)) & 1u ? 1.0 : 0.0;
  }
  return 0;
}

int rb200_trie_leaf_docs(const rb200_trie* tr, int64_t leaf, const int64_t** docs, int64_t* n) {
  RB_REQUIRE(tr && docs && n, "null argument");
  RB_REQUIRE(leaf >= 0 && leaf < tr->U, "leaf %lld outside [0, %lld)", (long long)leaf, (long long)tr->U);
  *docs = tr->leaf_docs.data() + tr->leaf_ptr[leaf];
  *n = tr->leaf_ptr[leaf + 1] - tr->leaf_ptr[leaf];
  return 0;
}

int rb200_trie_find_leaf(const rb200_trie* tr, const int32_t* code, int64_t* leaf) {
  RB_REQUIRE(tr && code && leaf, "null argument");
  const TrieView tv = tr->host_view();
  TrieState s = rb::trie_root(tv);
  for (int t = 0; t < tr->L; ++t) s = rb::trie_child(tv, s, t, code[t]);
  *leaf = (s.hi - s.lo == 1) ? s.lo : -1;
  return 0;
}

int rb200_trie_upload(rb200_trie* tr, int device) {
  RB_REQUIRE(tr, "null argument");
  if (tr->tables_on(device) != nullptr) {
    tr->device = device;
    return 0;
  }
  int ndev = 0;
  RB_CUDA(cudaGetDeviceCount(&ndev));
  RB_REQUIRE(device >= 0 && device < ndev, "device %d not present (%d devices)", device, ndev);
  int prev = 0;
  RB_CUDA(cudaGetDevice(&prev));
  RB_CUDA(cudaSetDevice(device));
  rb200_trie::DevTables t;
  t.device = device;
  auto up = [&](void** dst, const void* src, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(dst, bytes ? bytes : 4);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  };
  cudaError_t err = up(&t.codes, tr->codes.data(), tr->codes.size());
  if (err == cudaSuccess) err = up((void**)&t.node_bitmap, tr->node_bitmap.data(), 4 * tr->node_bitmap.size());
  if (err == cudaSuccess) err = up((void**)&t.node_child_ptr, tr->node_child_ptr.data(), 4 * tr->node_child_ptr.size());
  if (err == cudaSuccess) err = up((void**)&t.child_lo, tr->child_lo.data(), 4 * tr->child_lo.size());
  if (err == cudaSuccess) err = up((void**)&t.child_node, tr->child_node.data(), 4 * tr->child_node.size());
  if (err != cudaSuccess) {           // give back the partial copy
    cudaFree(t.codes); cudaFree(t.node_bitmap); cudaFree(t.node_child_ptr); cudaFree(t.child_lo); cudaFree(t.child_node);
    cudaSetDevice(prev);
    return rb::fail(RB200_ERR_CUDA, "rb200_trie_upload to device %d: %s", device, cudaGetErrorString(err));
  }
  tr->dev.push_back(t);
  tr->device = device;
  RB_CUDA(cudaSetDevice(prev));
  return 0;
}

int rb200_trie_leaf_expand(rb200_trie* tr, const int32_t* leaf_ranges, int64_t n, int k, int64_t* docs, int32_t* counts,
                           void* stream) {
  RB_REQUIRE(tr && leaf_ranges && docs && counts, "null argument");
  RB_REQUIRE(k >= 1 && k <= 4096, "max_docs_per_row=%d outside [1, 4096]", k);
  if (tr->dev.empty()) return rb::fail(RB200_ERR_STATE, "trie not uploaded: call rb200_trie_upload first");
  int cur = 0;
  RB_CUDA(cudaGetDevice(&cur));                    // the buffers belong to the calling thread's current device
  rb200_trie::DevTables* t = tr->tables_on(cur);
  if (t == nullptr)
    return rb::fail(RB200_ERR_STATE, "trie not uploaded to device %d: call rb200_trie_upload first", cur);
  if (n == 0) return 0;
  if (t->leaf_ptr == nullptr) {                    // first use: the leaf -> documents CSR goes to HBM as int32
    RB_REQUIRE(tr->n_docs < 0x7fffffff, "too many documents for the device leaf table");
    std::vector<int32_t> p32(tr->leaf_ptr.begin(), tr->leaf_ptr.end()), d32(tr->leaf_docs.begin(), tr->leaf_docs.end());
    RB_CUDA(cudaMalloc((void**)&t->leaf_ptr, p32.size() * 4));
    RB_CUDA(cudaMalloc((void**)&t->leaf_docs, d32.size() * 4 + 4));
    RB_CUDA(cudaMemcpy(t->leaf_ptr, p32.data(), p32.size() * 4, cudaMemcpyHostToDevice));
    RB_CUDA(cudaMemcpy(t->leaf_docs, d32.data(), d32.size() * 4, cudaMemcpyHostToDevice));
  }
  const int warps = 4;
  leaf_expand_kernel<<<rb::ceil_div(n, warps), warps * 32, 0, (cudaStream_t)stream>>>(
      t->leaf_ptr, t->leaf_docs, leaf_ranges, n, k, docs, counts);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

int rb200_trie_mask_device(const rb200_trie* tr, const int64_t* ids, int64_t R, int T, double* mask, void* stream) {
  RB_REQUIRE(tr && ids && mask, "null argument");
  if (tr->dev.empty()) return rb::fail(RB200_ERR_STATE, "trie not uploaded: call rb200_trie_upload first");
  int cur = 0;
  RB_CUDA(cudaGetDevice(&cur));
  if (tr->tables_on(cur) == nullptr)
    return rb::fail(RB200_ERR_STATE, "trie not uploaded to device %d: call rb200_trie_upload first", cur);
  RB_REQUIRE(T >= 1 && T <= tr->L, "prefix length T=%d outside [1, L=%d]", T, tr->L);
  if (R == 0) return 0;
  const int warps = 4;
  trie_mask_kernel<<<rb::ceil_div(R, warps), warps * 32, warps * tr->words * 4, (cudaStream_t)stream>>>(
      tr->device_view(cur), ids, R, T, mask);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

}  // extern "C"
