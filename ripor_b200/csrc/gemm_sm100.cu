// tcgen05 + TMA GEMM family for sm_100a:  C[M,N] = A[M,K] * W[N,K]^T  with fp32 accumulation in TMEM.
//
// Both operands are K-major and arrive already split into precision planes (act.cuh). One CTA computes a
// 128 x BN output tile, warp-specialised:
//   warp 0      TMA producer   cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 128 bytes along K) into a
//                              multi-stage shared-memory ring, completion on mbarriers
//   warp 1      MMA issuer     one elected lane issues tcgen05.mma (kind::tf32 or kind::f16) into a TMEM
//                              accumulator; NTERMS = 3 issues the error-compensated triple
//                              A_lo*W_hi + A_hi*W_lo + A_hi*W_hi per k-slice (fp32-grade products out of
//                              tf32 / bf16 tensor-core passes); tcgen05.commit releases the smem stage
//   warps 2..5  epilogue       tcgen05.ld (32 lanes x 32 columns per warp) -> registers -> fused epilogue
//                              (store / residual add / ReLU + re-split for the next GEMM) -> global
// Replaces the cuBLAS SGEMM calls behind every nn.Linear of the reference's T5 stack
// (t5_pretrainer/modeling/t5_generative_retriever.py:358-366,403-416 and get_lm_logits :250-262).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "kernels.h"

namespace rb {
namespace {

constexpr int BM = 128;
constexpr int SWIZZLE_BYTES = 128;       // one TMA box / UMMA swizzle atom along K
constexpr int kThreads = 192;
constexpr uint32_t kSmemBudget = 227 * 1024;

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  const long long t0 = clock64();
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000LL) __trap();   // a lost arrival must fail the launch, not hang the GPU
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
template <int KIND_TF32>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                       uint32_t accumulate) {
  if (KIND_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), LBO in [16,30) (unused for swizzled K-major, 1), SBO = 1024 B >> 4 in
// [32,46) (8 rows x 128 B between core-matrix groups), version 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

// cute::UMMA::InstrDescriptor: c_format F32 = 1 at [4,6); a/b format at [7,10)/[10,13) (BF16 = 1,
// TF32 = 2); K-major A and B; N >> 3 at [17,23); M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(BM >> 4) << 24);
}

template <int ELEM_BYTES, int NTERMS, int BN>
struct Cfg {
  static constexpr int PLANES = NTERMS == 3 ? 2 : 1;
  static constexpr int BK = SWIZZLE_BYTES / ELEM_BYTES;        // K elements per stage
  static constexpr int UMMA_K = 32 / ELEM_BYTES;               // K elements per tcgen05.mma
  static constexpr uint32_t A_TILE = BM * SWIZZLE_BYTES;       // 16 KB
  static constexpr uint32_t W_TILE = BN * SWIZZLE_BYTES;
  static constexpr uint32_t STAGE = PLANES * (A_TILE + W_TILE);
  static constexpr int STAGES_RAW = (kSmemBudget - 2048) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr uint32_t SMEM = STAGES * STAGE + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int ELEM_BYTES, int NTERMS, int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                  float* __restrict__ C, int64_t ldc, ActOut act, int M, int N, int K, int a_plane_rows,
                  int w_plane_rows, int epilogue, uint32_t idesc, float out_scale) {
  using cfg = Cfg<ELEM_BYTES, NTERMS, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + cfg::STAGES * cfg::STAGE);
  uint64_t* empty_bar = full_bar + cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + cfg::STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int num_kb = (K + cfg::BK - 1) / cfg::BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM allocation is warp-collective; the same warp frees it at the end
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % cfg::STAGES;
        const uint32_t ph = (kb / cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], cfg::STAGE);
        uint8_t* st = smem + s * cfg::STAGE;
        const int k0 = kb * cfg::BK;
#pragma unroll
        for (int p = 0; p < cfg::PLANES; ++p)
          tma_load_2d(&tmA, &full_bar[s], st + p * cfg::A_TILE, k0, m0 + p * a_plane_rows);
#pragma unroll
        for (int p = 0; p < cfg::PLANES; ++p)
          tma_load_2d(&tmW, &full_bar[s], st + cfg::PLANES * cfg::A_TILE + p * cfg::W_TILE, k0, n0 + p * w_plane_rows);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (single elected lane) =====
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % cfg::STAGES;
        const uint32_t ph = (kb / cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + s * cfg::STAGE);
        const uint32_t w_base = a_base + cfg::PLANES * cfg::A_TILE;
#pragma unroll
        for (int term = 0; term < NTERMS; ++term) {
          // NTERMS == 3: (A_lo, W_hi), (A_hi, W_lo), (A_hi, W_hi): small corrections first
          const int ap = (NTERMS == 3 && term == 0) ? 1 : 0;
          const int wp = (NTERMS == 3 && term == 1) ? 1 : 0;
          const uint64_t adesc = make_smem_desc(a_base + ap * cfg::A_TILE);
          const uint64_t wdesc = make_smem_desc(w_base + wp * cfg::W_TILE);
#pragma unroll
          for (int k = 0; k < cfg::BK / cfg::UMMA_K; ++k) {
            // advancing K inside the 128-byte swizzle atom = +32 bytes on the start-address field
            tc_mma<ELEM_BYTES == 4>(tmem_base, adesc + (uint64_t)(k * 2), wdesc + (uint64_t)(k * 2), idesc,
                                    (uint32_t)((kb | term | k) != 0));
          }
        }
        tc_commit(&empty_bar[s]);    // frees the smem stage once these MMAs have read it
      }
      tc_commit(tmem_full_bar);      // accumulator complete
    }
  } else {
    // ===== epilogue: warp w owns TMEM lanes [32*(w%4), +32) = tile rows =====
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int m = m0 + q * 32 + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
      if (m < M) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int n = n0 + c0 + j;
          if (n >= N) break;
          float4 v = make_float4(__uint_as_float(r[j]) * out_scale, __uint_as_float(r[j + 1]) * out_scale,
                                 __uint_as_float(r[j + 2]) * out_scale, __uint_as_float(r[j + 3]) * out_scale);
          if (epilogue == EPI_RELU_ACT) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
            act_store4(act, (int64_t)m * N + n, v);
          } else {
            float4* dst = reinterpret_cast<float4*>(C + (int64_t)m * ldc + n);
            if (epilogue == EPI_RESIDUAL) {
              const float4 o = *dst;
              v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
            }
            *dst = v;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps (cached), dispatch
// ------------------------------------------------------------------------------------------------
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

template <int ELEM_BYTES, int NTERMS, int BN>
int launch_cfg(const GemmArgs& g, cudaStream_t s) {
  using cfg = Cfg<ELEM_BYTES, NTERMS, BN>;
  static_assert(cfg::STAGES >= 2, "need at least a double-buffered pipeline");
  auto kern = gemm_sm100_kernel<ELEM_BYTES, NTERMS, BN>;
  static bool attr_set = false;
  if (!attr_set) {
    RB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM));
    attr_set = true;
  }
  RB_REQUIRE(g.a_plane % g.K == 0, "A plane distance must be a whole number of rows");
  const int64_t a_plane_rows = cfg::PLANES == 2 ? g.a_plane / g.K : 0;
  const int64_t w_plane_rows = cfg::PLANES == 2 ? g.w_plane / g.K : 0;
  const int64_t a_rows = cfg::PLANES == 2 ? a_plane_rows + g.M : g.M;
  const int64_t w_rows = cfg::PLANES == 2 ? w_plane_rows + g.N : g.N;
  CUtensorMap tmA, tmW;
  RB_TRY(get_tensor_map(g.A, a_rows, g.K, BM, ELEM_BYTES, &tmA));
  RB_TRY(get_tensor_map(g.W, w_rows, g.K, BN, ELEM_BYTES, &tmW));
  dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM));
  kern<<<grid, kThreads, cfg::SMEM, s>>>(tmA, tmW, g.C, g.ldc, g.act, (int)g.M, (int)g.N, (int)g.K,
                                         (int)a_plane_rows, (int)w_plane_rows, g.epilogue,
                                         make_idesc(ELEM_BYTES == 4 ? 2 : (prec_is_fp16(g.mode) ? 0 : 1), BN),
                                         g.out_scale);
  RB_CUDA(cudaGetLastError());
  rb::launch_count()++;
  return 0;
}

template <int ELEM_BYTES, int NTERMS>
int launch_bn(const GemmArgs& g, cudaStream_t s) {
  // 128-wide tiles when they still give every SM work, else 64-wide to raise the CTA count
  const int64_t tiles128 = (int64_t)ceil_div(g.M, BM) * ceil_div(g.N, 128);
  if (g.N > 64 && tiles128 >= 120) return launch_cfg<ELEM_BYTES, NTERMS, 128>(g, s);
  return launch_cfg<ELEM_BYTES, NTERMS, 64>(g, s);
}

}  // namespace

// shared with gemm_sm100_2cta.cu
int tensor_map_2d(const void* ptr, int64_t rows, int64_t k, int box_rows, int elem, CUtensorMap* out, int64_t ld) {
  return get_tensor_map(ptr, rows, k, box_rows, elem, out, ld);
}

int launch_gemm_sm100_2cta(const GemmArgs& g, cudaStream_t s);

int launch_gemm_sm100(const GemmArgs& g, cudaStream_t s) {
  // RB200_GEMM=1cta keeps the single-CTA 128xBN kernel; default is the cta_group::2 pair kernel
  static const bool use_pair = []() {
    const char* e = getenv("RB200_GEMM");
    return !(e && strcmp(e, "1cta") == 0);
  }();
  if (use_pair) return launch_gemm_sm100_2cta(g, s);
  RB_REQUIRE(g.epilogue != EPI_RESID_NORM && !g.nf.scaled, "the 1-CTA GEMM has no NormFold epilogues");
  const int elem = prec_elem_bytes(g.mode);
  RB_REQUIRE(g.K % (16 / elem) == 0, "K=%lld must be a multiple of %d for TMA", (long long)g.K, 16 / elem);
  RB_REQUIRE(g.N % 4 == 0, "N=%lld must be a multiple of 4", (long long)g.N);
  RB_REQUIRE(g.epilogue == EPI_RELU_ACT || g.ldc % 4 == 0, "ldc must be a multiple of 4");
  if (g.M == 0 || g.N == 0) return 0;
  switch (g.mode) {
    case RB200_PREC_TF32X3: return launch_bn<4, 3>(g, s);
    case RB200_PREC_BF16X3: return launch_bn<2, 3>(g, s);
    case RB200_PREC_TF32: return launch_bn<4, 1>(g, s);
    case RB200_PREC_BF16: return launch_bn<2, 1>(g, s);
    case RB200_PREC_FP16X3: return launch_bn<2, 3>(g, s);
    default: return fail(RB200_ERR_INVALID, "precision %d has no tensor-core GEMM", g.mode);
  }
}

}  // namespace rb
