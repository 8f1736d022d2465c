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
//
// The kernel is PERSISTENT: at most one cluster per SM pair, each walking tiles cluster_id, cluster_id + n, ...
// The shared-memory ring keeps streaming across tile boundaries and the TMEM accumulator is double-buffered
// (2 x BN columns), so the epilogue of tile i overlaps the MMAs of tile i+1 and the per-tile launch / prologue /
// drain cost of the one-tile-per-CTA version (about 5 us on a 10-25 us tile at these shapes) is paid once.
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
constexpr int kStgLd = 36;                               // padded row of the epilogue transpose tile (floats)
constexpr uint32_t kStgBytes = 4 * 32 * kStgLd * 4;      // one 32 x 36 fp32 tile per epilogue warp
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
// arrive on the barrier at the same offset in the leader CTA (rank 0) of the pair
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .b32 r;\n\t"
      "mapa.shared::cluster.u32 r, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t}"
      ::"r"(smem_u32(bar))
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
  static constexpr int STAGES_RAW = (kSmemBudget - 2048 - kStgBytes) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr uint32_t SMEM = STAGES * STAGE + 1024 + 256 + kStgBytes;
  static constexpr int TMEM_COLS = 2 * BN;                           // double-buffered accumulator
};

template <int ELEM_BYTES, int NTERMS, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_sm100_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       float* __restrict__ C, int64_t ldc, ActOut act, int M, int N, int K, int a_plane_rows,
                       int w_plane_rows, int epilogue, uint32_t idesc, float out_scale, int num_n_tiles,
                       int num_tiles) {
  using cfg = Cfg2<ELEM_BYTES, NTERMS, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + cfg::STAGES * cfg::STAGE);
  uint64_t* empty_bar = full_bar + cfg::STAGES;
  uint64_t* tmem_full_bar = empty_bar + cfg::STAGES;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;         // [2], only the leader's copies are used
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_kb = (K + cfg::BK - 1) / cfg::BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    for (int s = 0; s < cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 8);      // 4 epilogue warps x 2 CTAs
    }
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
  // everything above overlapped the tail of the previous kernel (PDL); operands are read only from here on
  pdl_trigger();
  pdl_wait();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int m0 = ((tile / num_n_tiles) * 2 + (int)rank) * BM;       // this CTA's 128 rows of the pair tile
        const int w0 = (tile % num_n_tiles) * BN + (int)rank * (BN / 2);  // this CTA's half of the W rows
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % cfg::STAGES;
          const uint32_t ph = (it / cfg::STAGES) & 1;
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
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      int it = 0, lt = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++lt) {
        const int buf = lt & 1;
        mbar_wait(&tmem_empty_bar[buf], ((lt >> 1) & 1) ^ 1);     // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % cfg::STAGES;
          const uint32_t ph = (it / cfg::STAGES) & 1;
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
              tc_mma_pair<ELEM_BYTES == 4>(tmem_d, adesc + (uint64_t)(k * 2), wdesc + (uint64_t)(k * 2), idesc,
                                           (uint32_t)((kb | term | k) != 0));
          }
          tc_commit_pair(&empty_bar[s]);
        }
        tc_commit_pair(&tmem_full_bar[buf]);
      }
    }
  } else {
    const int q = warp & 3;
    float* stg = reinterpret_cast<float*>(smem + cfg::STAGES * cfg::STAGE + 256) + (warp - 2) * (32 * kStgLd);
    int lt = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++lt) {
    const int buf = lt & 1;
    const int m0 = ((tile / num_n_tiles) * 2 + (int)rank) * BM;
    const int n0 = (tile % num_n_tiles) * BN;
    // Residual epilogue: the old values of C are an INPUT, so fetch them while the MMAs are still running
    // (chunk 0 before waiting on the accumulator, chunk i+1 while chunk i is being transposed). Loads are issued
    // before the stores of the previous chunk in program order: the compiler may not hoist them itself (aliasing).
    auto load_residual = [&](int c0, float4 (&res)[8]) {
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        const int m = m0 + q * 32 + itr * 4 + (lane >> 3), n = n0 + c0 + (lane & 7) * 4;
        res[itr] = (m < M && n < N) ? *reinterpret_cast<const float4*>(C + (int64_t)m * ldc + n)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    float4 res[8];
    if (epilogue == EPI_RESIDUAL) load_residual(0, res);
    mbar_wait(&tmem_full_bar[buf], (lt >> 1) & 1);
    tc_fence_after();
    // TMEM gives each lane one ROW (32 consecutive columns). Storing that way makes every warp store touch 32
    // different lines; transposing the 32x32 chunk through a padded per-warp shared-memory tile lets 8 lanes
    // write 128 contiguous bytes of one row (4 rows per instruction), which cuts the LSU wavefronts 8x.
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + c0), r);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        *reinterpret_cast<float4*>(&stg[lane * kStgLd + j * 4]) =
            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                        __uint_as_float(r[4 * j + 3]));
      __syncwarp();
      float4 vals[8];
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        float4 v = *reinterpret_cast<const float4*>(&stg[(itr * 4 + (lane >> 3)) * kStgLd + (lane & 7) * 4]);
        v.x *= out_scale; v.y *= out_scale; v.z *= out_scale; v.w *= out_scale;
        if (epilogue == EPI_RESIDUAL) { v.x += res[itr].x; v.y += res[itr].y; v.z += res[itr].z; v.w += res[itr].w; }
        if (epilogue == EPI_RELU_ACT) {
          v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        vals[itr] = v;
      }
      if (epilogue == EPI_RESIDUAL && c0 + 32 < BN) load_residual(c0 + 32, res);
#pragma unroll
      for (int itr = 0; itr < 8; ++itr) {
        const int m = m0 + q * 32 + itr * 4 + (lane >> 3), n = n0 + c0 + (lane & 7) * 4;
        if (m < M && n < N) {
          if (epilogue == EPI_RELU_ACT) act_store4(act, (int64_t)m * N + n, vals[itr]);
          else *reinterpret_cast<float4*>(C + (int64_t)m * ldc + n) = vals[itr];
        }
      }
      __syncwarp();
    }
    // this warp is done reading the accumulator: tell the leader's MMA warp it may be overwritten
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);
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
  const int num_n_tiles = ceil_div(g.N, BN);
  const int num_tiles = ceil_div(g.M, 2 * BM) * num_n_tiles;
  static int sm_pairs = 0;
  if (sm_pairs == 0) {
    int dev = 0, sms = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sm_pairs = sms / 2;
  }
  dim3 grid(2 * (num_tiles < sm_pairs ? num_tiles : sm_pairs));
  const int fmt = ELEM_BYTES == 4 ? 2 : (prec_is_fp16(g.mode) ? 0 : 1);
  RB_CUDA(launch_pdl(kern, grid, dim3(kThreads), cfg::SMEM, s, tmA, tmW, g.C, g.ldc, g.act, (int)g.M, (int)g.N,
                     (int)g.K, (int)a_plane_rows, (int)w_plane_rows, g.epilogue, make_idesc_pair(fmt, BN), g.out_scale,
                     num_n_tiles, num_tiles));
  launch_count()++;
  return 0;
}

template <int ELEM_BYTES, int NTERMS>
int launch_bn2(const GemmArgs& g, cudaStream_t s) {
  // Persistent schedule: every SM pair walks ceil(tiles / pairs) tiles whose duration grows with BN. Pick the
  // tile width with the smaller (rounds x BN); ties go to 256 (fewer operand bytes per output).
  static int sm_pairs = 0;
  if (sm_pairs == 0) {
    int dev = 0, sms = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    sm_pairs = sms / 2;
  }
  const int64_t m_pairs = ceil_div(g.M, 2 * BM);
  const int64_t cost256 = (int64_t)ceil_div(m_pairs * ceil_div(g.N, 256), sm_pairs) * 256;
  const int64_t cost128 = (int64_t)ceil_div(m_pairs * ceil_div(g.N, 128), sm_pairs) * 128;
  if (g.N >= 256 && cost256 <= cost128) return launch_cfg2<ELEM_BYTES, NTERMS, 256>(g, s);
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
