// Cached cuTensorMap descriptors for the tcgen05 / TMA GEMM family (gemm_sm100_2cta.cu) and the dispatch into it.
//
// The engine launches ~450 GEMMs per search over a few dozen distinct (buffer, shape) pairs; encoding a tensor map
// costs microseconds of driver time, so they are built once per (pointer, rows, k, ld, box, element size).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "kernels.h"

namespace rb {
namespace {

constexpr int SWIZZLE_BYTES = 128;       // one TMA box / UMMA swizzle atom along K

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  int64_t rows, k, ld;
  int box_rows, elem;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && rows == o.rows && k == o.k && ld == o.ld && box_rows == o.box_rows && elem == o.elem;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<int64_t>()(k.rows);
    h = h * 1000003u ^ std::hash<int64_t>()(k.k);
    h = h * 1000003u ^ std::hash<int64_t>()(k.ld);
    h = h * 1000003u ^ (size_t)(k.box_rows * 8 + k.elem);
    return h;
  }
};

// 2-D row-major tensor [rows, k] with row stride `ld` elements; boxes of box_rows x 128 bytes, SWIZZLE_128B.
// Used for the operand loads (ld == k) and for the epilogue's TMA stores / reduce-adds of output tiles.
int get_tensor_map(const void* ptr, int64_t rows, int64_t k, int box_rows, int elem, CUtensorMap* out,
                   int64_t ld = 0) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  if (ld == 0) ld = k;
  const MapKey key{ptr, rows, k, ld, box_rows, elem};
  {
    std::lock_guard<std::mutex> lock(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  auto encode = get_encode_fn();
  if (!encode) return fail(RB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * elem};
  const cuuint32_t box[2] = {(cuuint32_t)(SWIZZLE_BYTES / elem), (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  const CUresult r = encode(&m, elem == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                            const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(RB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for ptr=%p rows=%lld k=%lld ld=%lld box_rows=%d elem=%d",
                (int)r, ptr, (long long)rows, (long long)k, (long long)ld, box_rows, elem);
  {
    std::lock_guard<std::mutex> lock(mu);
    if (cache.size() >= 4096) cache.clear();   // callers that keep passing fresh buffers must not grow it forever
    cache[key] = m;
  }
  *out = m;
  return 0;
}

}  // namespace

int tensor_map_2d(const void* ptr, int64_t rows, int64_t k, int box_rows, int elem, CUtensorMap* out, int64_t ld) {
  return get_tensor_map(ptr, rows, k, box_rows, elem, out, ld);
}

int launch_gemm_sm100_2cta(const GemmArgs& g, cudaStream_t s);

// every tensor-core precision runs on the cta_group::2 pair kernel (the round-1 single-CTA kernel is gone)
int launch_gemm_sm100(const GemmArgs& g, cudaStream_t s) { return launch_gemm_sm100_2cta(g, s); }

}  // namespace rb
