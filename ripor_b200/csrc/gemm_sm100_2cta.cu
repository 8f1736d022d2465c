// tcgen05 cta_group::2 GEMM: a CTA pair (thread-block cluster of 2, one SM pair) computes a 256 x BN tile.
//
// ncu on the 1-CTA kernel (profiles/r01_*) shows the GEMMs of this path are bound by operand delivery from
// L2 to shared memory (~10 TB/s over the whole chip), not by the tensor pipe: a 128x128 tile re-reads
// (128+128) operand rows per 128x128 outputs. Pairing two SMs halves that: each CTA stages its own 128 rows of A
// and only HALF of the W tile (BN/2 rows); the leader CTA issues tcgen05.mma.cta_group::2 (M = 256) which reads
// both CTAs' shared memory and writes both CTAs' TMEM; (256+BN) rows feed 256 x BN outputs.
//
//   both CTAs   warp 0: TMA producer (cp.async.bulk.tensor.2d.cta_group::2, completion signalled on the
//                       LEADER's full barrier), waits on its own empty barrier
//   leader CTA  warp 1: MMA issuer; tcgen05.commit.cta_group::2 ... multicast::cluster releases the stage in
//                       both CTAs and finally publishes the accumulator to both epilogues
//   both CTAs   warps 2..5: epilogue for their own 128 rows (tcgen05.ld -> registers -> fused epilogue)
#include <cuda.h>
#include <cudaTypedefs.h>

#include "kernels.h"

namespace rb {

int tensor_map_2d(const void* ptr, int64_t rows, int64_t k, int box_rows, int elem, CUtensorMap* out);

namespace {

constexpr int BM = 128;                  // rows per CTA (256 per pair)
constexpr int SWIZZLE_BYTES = 128;
constexpr int kThreads = 192;
constexpr uint32_t kSmemBudget = 227 * 1024;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // shared::cluster address of the same offset in the even (leader) CTA

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
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
// 2-SM TMA load: data lands in THIS CTA's shared memory, the transaction bytes are reported to the barrier at
// the same offset in the leader CTA.
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// arrive (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"((uint16_t)3)
      : "memory");
}
template <int KIND_TF32>
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  if (KIND_TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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
// K-major SWIZZLE_128B descriptor (see gemm_sm100.cu)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// instruction descriptor with M = 256 (pair), N = BN
__host__ __device__ constexpr uint32_t make_idesc_pair(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

template <int ELEM_BYTES, int NTERMS, int BN>
struct Cfg2 {
  static constexpr int PLANES = NTERMS == 3 ? 2 : 1;
  static constexpr int BK = SWIZZLE_BYTES / ELEM_BYTES;
  static constexpr int UMMA_K = 32 / ELEM_BYTES;
  static constexpr uint32_t A_TILE = BM * SWIZZLE_BYTES;            // this CTA's 128 rows of A
  static constexpr uint32_t W_TILE = (BN / 2) * SWIZZLE_BYTES;      // this CTA's half of the W tile
  static constexpr uint32_t STAGE = PLANES * (A_TILE + W_TILE);
  static constexpr int STAGES_RAW = (kSmemBudget - 2048) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr uint32_t SMEM = STAGES * STAGE + 1024 + 256;
  static constexpr int TMEM_COLS = BN < 32 ? 32 : BN;
};

template <int ELEM_BYTES, int NTERMS, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_sm100_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       float* __restrict__ C, int64_t ldc, ActOut act, int M, int N, int K, int a_plane_rows,
                       int w_plane_rows, int epilogue, uint32_t idesc, float out_scale) {
  using cfg = Cfg2<ELEM_BYTES, NTERMS, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + cfg::STAGES * cfg::STAGE);
  uint64_t* empty_bar = full_bar + cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + cfg::STAGES;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int m0 = blockIdx.x * BM;                       // this CTA's rows (pair = two consecutive blockIdx.x)
  const int n0 = blockIdx.y * BN;                       // the pair's columns
  const int w0 = n0 + (int)rank * (BN / 2);             // this CTA's half of the W rows
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
  if (warp == 1) {   // both CTAs' warp 1, same destination offset (cute::TMEM::Allocator2Sm contract)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();          // barriers of both CTAs initialised and TMEM allocated before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % cfg::STAGES;
        const uint32_t ph = (kb / cfg::STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (leader) mbar_expect_tx(&full_bar[s], 2 * cfg::STAGE);      // both CTAs' loads report here
        uint8_t* st = smem + s * cfg::STAGE;
        const int k0 = kb * cfg::BK;
#pragma unroll
        for (int p = 0; p < cfg::PLANES; ++p)
          tma_load_2d_2sm(&tmA, &full_bar[s], st + p * cfg::A_TILE, k0, m0 + p * a_plane_rows);
#pragma unroll
        for (int p = 0; p < cfg::PLANES; ++p)
          tma_load_2d_2sm(&tmW, &full_bar[s], st + cfg::PLANES * cfg::A_TILE + p * cfg::W_TILE, k0,
                          w0 + p * w_plane_rows);
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % cfg::STAGES;
        const uint32_t ph = (kb / cfg::STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_base = smem_u32(smem + s * cfg::STAGE);
        const uint32_t w_base = a_base + cfg::PLANES * cfg::A_TILE;
#pragma unroll
        for (int term = 0; term < NTERMS; ++term) {
          const int ap = (NTERMS == 3 && term == 0) ? 1 : 0;
          const int wp = (NTERMS == 3 && term == 1) ? 1 : 0;
          const uint64_t adesc = make_smem_desc(a_base + ap * cfg::A_TILE);
          const uint64_t wdesc = make_smem_desc(w_base + wp * cfg::W_TILE);
#pragma unroll
          for (int k = 0; k < cfg::BK / cfg::UMMA_K; ++k)
            tc_mma_pair<ELEM_BYTES == 4>(tmem_base, adesc + (uint64_t)(k * 2), wdesc + (uint64_t)(k * 2), idesc,
                                         (uint32_t)((kb | term | k) != 0));
        }
        tc_commit_pair(&empty_bar[s]);
      }
      tc_commit_pair(tmem_full_bar);
    }
  } else {
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
  cluster_sync_all();          // neither CTA may free TMEM or exit while the peer can still touch it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
  }
}

template <int ELEM_BYTES, int NTERMS, int BN>
int launch_cfg2(const GemmArgs& g, cudaStream_t s) {
  using cfg = Cfg2<ELEM_BYTES, NTERMS, BN>;
  static_assert(cfg::STAGES >= 2, "need at least a double-buffered pipeline");
  auto kern = gemm_sm100_2cta_kernel<ELEM_BYTES, NTERMS, BN>;
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
  RB_TRY(tensor_map_2d(g.A, a_rows, g.K, BM, ELEM_BYTES, &tmA));
  RB_TRY(tensor_map_2d(g.W, w_rows, g.K, BN / 2, ELEM_BYTES, &tmW));
  dim3 grid(2 * ceil_div(g.M, 2 * BM), ceil_div(g.N, BN));
  const int fmt = ELEM_BYTES == 4 ? 2 : (prec_is_fp16(g.mode) ? 0 : 1);
  kern<<<grid, kThreads, cfg::SMEM, s>>>(tmA, tmW, g.C, g.ldc, g.act, (int)g.M, (int)g.N, (int)g.K,
                                         (int)a_plane_rows, (int)w_plane_rows, g.epilogue, make_idesc_pair(fmt, BN),
                                         g.out_scale);
  RB_CUDA(cudaGetLastError());
  launch_count()++;
  return 0;
}

template <int ELEM_BYTES, int NTERMS>
int launch_bn2(const GemmArgs& g, cudaStream_t s) {
  // 256x256 pair tiles when they still occupy most SMs, else 256x128
  const int64_t ctas256 = 2 * (int64_t)ceil_div(g.M, 2 * BM) * ceil_div(g.N, 256);
  if (g.N >= 256 && ctas256 >= 100) return launch_cfg2<ELEM_BYTES, NTERMS, 256>(g, s);
  return launch_cfg2<ELEM_BYTES, NTERMS, 128>(g, s);
}

}  // namespace

int launch_gemm_sm100_2cta(const GemmArgs& g, cudaStream_t s) {
  const int elem = prec_elem_bytes(g.mode);
  RB_REQUIRE(g.K % (16 / elem) == 0, "K=%lld must be a multiple of %d for TMA", (long long)g.K, 16 / elem);
  RB_REQUIRE(g.N % 4 == 0, "N=%lld must be a multiple of 4", (long long)g.N);
  RB_REQUIRE(g.epilogue == EPI_RELU_ACT || g.ldc % 4 == 0, "ldc must be a multiple of 4");
  if (g.M == 0 || g.N == 0) return 0;
  switch (g.mode) {
    case RB200_PREC_TF32X3: return launch_bn2<4, 3>(g, s);
    case RB200_PREC_BF16X3: return launch_bn2<2, 3>(g, s);
    case RB200_PREC_FP16X3: return launch_bn2<2, 3>(g, s);
    case RB200_PREC_TF32: return launch_bn2<4, 1>(g, s);
    case RB200_PREC_BF16: return launch_bn2<2, 1>(g, s);
    default: return fail(RB200_ERR_INVALID, "precision %d has no tensor-core GEMM", g.mode);
  }
}

}  // namespace rb
