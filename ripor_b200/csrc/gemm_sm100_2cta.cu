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
#include <cstdlib>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "kernels.h"

namespace rb {

static unsigned long long* g_gemm_trace = nullptr;     // device buffer [grid][8] or null
void set_gemm_trace(unsigned long long* dev) { g_gemm_trace = dev; }

int tensor_map_2d(const void* ptr, int64_t rows, int64_t k, int box_rows, int elem, CUtensorMap* out, int64_t ld);

namespace {

constexpr int BM = 128;                  // rows per CTA (256 per pair)
constexpr int SWIZZLE_BYTES = 128;
constexpr int kThreads = 64 + 8 * 32;   // TMA warp, MMA warp, 8 epilogue warps
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

// optional phase trace (tools/gemm_bench.py --trace): %globaltimer stamps per CTA, 8 slots each
__device__ __forceinline__ void trace_stamp(unsigned long long* trace, int slot) {
  if (trace) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    trace[blockIdx.x * 8 + slot] = t;
  }
}

constexpr int kEpiWarps = 8;                             // two per TMEM lane quarter, each owning half of the columns
constexpr uint32_t kEpiTile = 32 * SWIZZLE_BYTES;        // one 32-row x 128-byte staging tile per epilogue warp

template <int ELEM_BYTES, int NTERMS, int BN>
struct Cfg2 {
  static constexpr int PLANES = NTERMS == 3 ? 2 : 1;
  static constexpr int BK = SWIZZLE_BYTES / ELEM_BYTES;
  static constexpr int UMMA_K = 32 / ELEM_BYTES;
  static constexpr uint32_t A_TILE = BM * SWIZZLE_BYTES;            // this CTA's 128 rows of A
  static constexpr uint32_t W_TILE = (BN / 2) * SWIZZLE_BYTES;      // this CTA's half of the W tile
  static constexpr uint32_t STAGE = PLANES * (A_TILE + W_TILE);
  static constexpr uint32_t EPI_BYTES = kEpiWarps * kEpiTile;
  static constexpr int STAGES_RAW = (kSmemBudget - 2048 - EPI_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr uint32_t SMEM = STAGES * STAGE + EPI_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = 2 * BN;                           // double-buffered accumulator
};

// ---- epilogue helpers: registers -> swizzled staging tile -> TMA store / reduce-add -----------------------------
// A staging tile is 32 rows x 128 bytes in the SWIZZLE_128B layout the output tensor maps expect: the 16-byte chunk
// j of row r lives at r * 128 + ((j ^ (r & 7)) << 4). Lane = row, so the eight 16-byte stores of a lane are
// bank-conflict free, and the TMA engine turns the tile into full-line global writes (or, for the residual
// epilogue, into fp32 reduce-adds performed by L2: the old values of C never travel to the SM).
__device__ __forceinline__ void stage_row(uint8_t* tile, int lane, const uint32_t (&w)[32]) {
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<uint4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_uint4(w[4 * j], w[4 * j + 1], w[4 * j + 2], w[4 * j + 3]);
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1) : "memory");
}
// one tile of this warp: wait until the previous bulk store has read the staging tile, refill it, hand it to TMA
__device__ __forceinline__ void emit_tile(uint8_t* tile, int lane, const uint32_t (&w)[32], const CUtensorMap* map,
                                          int c0, int c1, bool reduce_add) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
  stage_row(tile, lane, w);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    if (reduce_add) tma_reduce_add_2d(map, tile, c0, c1);
    else tma_store_2d(map, tile, c0, c1);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

template <int ELEM_BYTES, int NTERMS, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_sm100_2cta_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                       const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
                       const __grid_constant__ CUtensorMap tmO2, const float* __restrict__ Cin, int64_t ldc,
                       NormFold nf, int* overflow, int is_fp16, int M, int N, int K, int a_plane_rows,
                       int w_plane_rows, int epilogue, uint32_t idesc, float out_scale, int num_n_tiles,
                       int num_tiles, unsigned long long* trace) {
  using cfg = Cfg2<ELEM_BYTES, NTERMS, BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + cfg::STAGES * cfg::STAGE;              // 1024-aligned: STAGE is a multiple of 1024
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + cfg::EPI_BYTES);
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
    trace_stamp(trace, 0);
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmW) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmO2) : "memory");
    for (int s = 0; s < cfg::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 2 * kEpiWarps);      // every epilogue warp of both CTAs
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

  // this CTA's share of one pipeline iteration `it` (tile-major, then k-block)
  auto load_w = [&](int tile, int kb, int s) {
    const int w0 = (tile % num_n_tiles) * BN + (int)rank * (BN / 2);      // this CTA's half of the W rows
    uint8_t* st = smem + s * cfg::STAGE;
#pragma unroll
    for (int p = 0; p < cfg::PLANES; ++p)
      tma_load_2d_2sm(&tmW, &full_bar[s], st + cfg::PLANES * cfg::A_TILE + p * cfg::W_TILE, kb * cfg::BK,
                      w0 + p * w_plane_rows);
  };
  auto load_a = [&](int tile, int kb, int s) {
    const int m0 = ((tile / num_n_tiles) * 2 + (int)rank) * BM;           // this CTA's 128 rows of the pair tile
    uint8_t* st = smem + s * cfg::STAGE;
#pragma unroll
    for (int p = 0; p < cfg::PLANES; ++p)
      tma_load_2d_2sm(&tmA, &full_bar[s], st + p * cfg::A_TILE, kb * cfg::BK, m0 + p * a_plane_rows);
  };
  // The weights do not depend on the previous kernel of the stream: their first pipeline stages are requested
  // BEFORE the grid dependency resolves (PDL), so the HBM latency of the first W tiles overlaps that kernel's tail.
  int prefetched = 0;
  if (warp == 0 && lane == 0) {
    const int my_tiles = (num_tiles - cluster_id + num_clusters - 1) / num_clusters;
    const int total_it = my_tiles * num_kb;
    prefetched = total_it < cfg::STAGES ? total_it : cfg::STAGES;
    for (int it = 0; it < prefetched; ++it) {
      if (leader) mbar_expect_tx(&full_bar[it], 2 * cfg::STAGE);          // both CTAs' A and W loads report here
      load_w(cluster_id + (it / num_kb) * num_clusters, it % num_kb, it);
    }
  }
  pdl_wait();                  // from here on the previous kernel's outputs (A, C) may be touched
  const uint32_t tmem_base = *tmem_base_slot;
  if (warp == 0 && lane == 0) trace_stamp(trace, 1);

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % cfg::STAGES;
          if (it >= prefetched) {
            const uint32_t ph = (it / cfg::STAGES) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (leader) mbar_expect_tx(&full_bar[s], 2 * cfg::STAGE);
            load_w(tile, kb, s);
          }
          load_a(tile, kb, s);
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
          if (it == 0) trace_stamp(trace, 2);
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
      trace_stamp(trace, 3);
    }
  } else {
    // ===== epilogue: warp -> (TMEM lane quarter q = warp % 4, column half ch); lane = row of the quarter =====
    const int q = warp & 3, ch = (warp - 2) >> 2;
    uint8_t* tile_s = epi_smem + (warp - 2) * kEpiTile;
    // 64-wide tiles (the latency-bound GEMMs of small batches): the 16-bit plane epilogues work on 64 columns at a
    // time, so one warp per lane quarter takes the whole width and its column-half partner only keeps the barrier
    // protocol going
    constexpr int HALF = BN >= 128 ? BN / 2 : BN;
    const bool epi_active = BN >= 128 || ch == 0;
    bool bad = false;                                  // fp16 planes: a value left the representable range
    int lt = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++lt) {
      const int buf = lt & 1;
      if (!epi_active) {
        mbar_wait(&tmem_full_bar[buf], (lt >> 1) & 1);
        if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);
        continue;
      }
      const int row0 = ((tile / num_n_tiles) * 2 + (int)rank) * BM + q * 32;
      const int col0 = (tile % num_n_tiles) * BN + ch * HALF;
      // ---- everything that does not need the accumulator happens while the MMAs are still running ----
      // NormFold: r^2 = mean(x^2) + eps from the per-row partial sums (lane = row)
      const int myrow = row0 + lane;
      float rowscale = out_scale, inv_ref = 1.0f;
      if (nf.ss_prev != nullptr) {
        float sp = 0.f, sc = 0.f;
        if (myrow < M) {
          for (int i = 0; i < nf.np; ++i) sp += nf.ss_prev[(int64_t)myrow * nf.np + i];
          if (nf.scaled)
            for (int i = 0; i < nf.np; ++i) sc += nf.ss_cur[(int64_t)myrow * nf.np + i];
        }
        const float r2_prev = fmaf(sp, nf.inv_d, nf.eps);
        inv_ref = 1.0f / sqrtf(r2_prev);
        if (nf.scaled) rowscale = out_scale * sqrtf(r2_prev / fmaf(sc, nf.inv_d, nf.eps));
      }
      // EPI_RESID_NORM: the old values of this lane's row for the first 64 columns (row-per-lane 128-byte reads)
      const float* xrow = Cin + (int64_t)(myrow < M ? myrow : 0) * ldc;
      auto load_x = [&](int cb, float4 (&xo)[8]) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          xo[j] = (myrow < M && cb + 4 * j < N) ? *reinterpret_cast<const float4*>(xrow + cb + 4 * j)
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      float4 xo[2][8];
      if (epilogue == EPI_RESID_NORM) {
        load_x(col0, xo[0]);
        load_x(col0 + 32, xo[1]);
      }
      mbar_wait(&tmem_full_bar[buf], (lt >> 1) & 1);
      tc_fence_after();
      if (warp == 2 && lane == 0) {
        if (lt == 0) trace_stamp(trace, 4);
        // the last accumulator of this CTA is complete: only epilogues remain, let the next kernel of the stream
        // start its prologue (PDL). Triggering earlier would park its CTAs on SMs this grid still needs.
        if (tile + num_clusters >= num_tiles) pdl_trigger();
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + ch * HALF);
      const bool rows_live = row0 < M;                 // warp-uniform; TMA clips partially covered boxes itself
      if (epilogue == EPI_RESID_NORM) {
        // C = C + acc; planes(C / r_prev) for the consuming GEMM; per-row partial sums of C^2 (NormFold)
#pragma unroll 1
        for (int c = 0; c < HALF; c += 64) {
          uint32_t xn[2][32];
          float ssq = 0.f;
#pragma unroll
          for (int hc = 0; hc < 2; ++hc) {
            uint32_t r[32];
            tmem_ld32(taddr + (uint32_t)(c + hc * 32), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float v0 = fmaf(__uint_as_float(r[4 * j + 0]), out_scale, xo[hc][j].x);
              const float v1 = fmaf(__uint_as_float(r[4 * j + 1]), out_scale, xo[hc][j].y);
              const float v2 = fmaf(__uint_as_float(r[4 * j + 2]), out_scale, xo[hc][j].z);
              const float v3 = fmaf(__uint_as_float(r[4 * j + 3]), out_scale, xo[hc][j].w);
              ssq += v0 * v0 + v1 * v1 + v2 * v2 + v3 * v3;
              xn[hc][4 * j + 0] = __float_as_uint(v0); xn[hc][4 * j + 1] = __float_as_uint(v1);
              xn[hc][4 * j + 2] = __float_as_uint(v2); xn[hc][4 * j + 3] = __float_as_uint(v3);
            }
            // this slot's old values are consumed: request the same 32 columns of the next 64-column chunk now, so
            // that they arrive while the tile stores and the plane conversion of this chunk run
            if (c + 64 < HALF) load_x(col0 + c + 64 + hc * 32, xo[hc]);
            // the new residual values leave through the staging tile as one TMA store per 32 columns (row-per-lane
            // 16-byte stores touch 32 rows per instruction: the same scattered-store pattern that held the tail
            // attention kernels back)
            if (rows_live && col0 + c + hc * 32 < N) emit_tile(tile_s, lane, xn[hc], &tmO0, col0 + c + hc * 32, row0, false);
          }
          if (myrow < M && col0 + c < N) nf.ss_out[(int64_t)myrow * nf.np + ((col0 + c) >> 6)] = ssq;
          // operand planes of x / r_prev for the next GEMM
          if (ELEM_BYTES == 4) {
#pragma unroll
            for (int hc = 0; hc < 2; ++hc) {
              uint32_t hi[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(xn[hc][j]) * inv_ref;
                const float h = round_tf32(v);
                hi[j] = __float_as_uint(h);
                xn[hc][j] = __float_as_uint(round_tf32(v - h));
              }
              if (rows_live && col0 + c + hc * 32 < N) {
                emit_tile(tile_s, lane, hi, &tmO1, col0 + c + hc * 32, row0, false);
                if (NTERMS == 3) emit_tile(tile_s, lane, xn[hc], &tmO2, col0 + c + hc * 32, row0, false);
              }
            }
          } else {
            uint32_t hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = __uint_as_float(j < 16 ? xn[0][2 * j] : xn[1][2 * j - 32]) * inv_ref;
              const float b = __uint_as_float(j < 16 ? xn[0][2 * j + 1] : xn[1][2 * j - 31]) * inv_ref;
              if (is_fp16) {
                bad |= !(fabsf(a) <= kFp16Limit && fabsf(b) <= kFp16Limit);
                const __half2 h = __floats2half2_rn(a, b);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
                hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                lo[j] = *reinterpret_cast<const uint32_t*>(&l);
              } else {
                const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
                const float2 hf = __bfloat1622float2(h);
                const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
                hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                lo[j] = *reinterpret_cast<const uint32_t*>(&l);
              }
            }
            if (rows_live && col0 + c < N) {
              emit_tile(tile_s, lane, hi, &tmO1, col0 + c, row0, false);
              if (NTERMS == 3) emit_tile(tile_s, lane, lo, &tmO2, col0 + c, row0, false);
            }
          }
        }
      } else if (epilogue != EPI_RELU_ACT && epilogue != EPI_PLANES) {
        // fp32 output: 32 columns = 128 bytes per staging row
#pragma unroll 1
        for (int c = 0; c < HALF; c += 32) {
          uint32_t r[32];
          tmem_ld32(taddr + (uint32_t)c, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * rowscale);
          if (rows_live && col0 + c < N) emit_tile(tile_s, lane, r, &tmO0, col0 + c, row0, epilogue == EPI_RESIDUAL);
        }
      } else if (ELEM_BYTES == 4) {
        // ReLU -> tf32 planes kept as fp32 words: 32 columns per tile, plane 0 then (x3 modes) plane 1
#pragma unroll 1
        for (int c = 0; c < HALF; c += 32) {
          uint32_t r[32], hi[32];
          tmem_ld32(taddr + (uint32_t)c, r);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = __uint_as_float(r[j]) * rowscale;
            const float v = epilogue == EPI_RELU_ACT ? fmaxf(x, 0.f) : x;
            const float h = round_tf32(v);
            hi[j] = __float_as_uint(h);
            r[j] = __float_as_uint(round_tf32(v - h));
          }
          if (rows_live && col0 + c < N) {
            emit_tile(tile_s, lane, hi, &tmO0, col0 + c, row0, false);
            if (NTERMS == 3) emit_tile(tile_s, lane, r, &tmO1, col0 + c, row0, false);
          }
        }
      } else {
        // ReLU -> 16-bit planes: 64 columns = 128 bytes per staging row
#pragma unroll 1
        for (int c = 0; c < HALF; c += 64) {
          uint32_t r0[32], r1[32], hi[32], lo[32];
          tmem_ld32(taddr + (uint32_t)c, r0);
          tmem_ld32(taddr + (uint32_t)(c + 32), r1);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(j < 16 ? r0[2 * j] : r1[2 * j - 32]) * rowscale;
            float b = __uint_as_float(j < 16 ? r0[2 * j + 1] : r1[2 * j - 31]) * rowscale;
            if (epilogue == EPI_RELU_ACT) { a = fmaxf(a, 0.f); b = fmaxf(b, 0.f); }
            if (is_fp16) {
              bad |= !(fabsf(a) <= kFp16Limit && fabsf(b) <= kFp16Limit);
              const __half2 h = __floats2half2_rn(a, b);
              const float2 hf = __half22float2(h);
              const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
              hi[j] = *reinterpret_cast<const uint32_t*>(&h);
              lo[j] = *reinterpret_cast<const uint32_t*>(&l);
            } else {
              const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
              const float2 hf = __bfloat1622float2(h);
              const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
              hi[j] = *reinterpret_cast<const uint32_t*>(&h);
              lo[j] = *reinterpret_cast<const uint32_t*>(&l);
            }
          }
          if (rows_live && col0 + c < N) {
            emit_tile(tile_s, lane, hi, &tmO0, col0 + c, row0, false);
            if (NTERMS == 3) emit_tile(tile_s, lane, lo, &tmO1, col0 + c, row0, false);
          }
        }
      }
      // this warp is done reading the accumulator: tell the leader's MMA warp it may be overwritten
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);
    }
    // the staging tile must outlive the bulk store's READ of it; the global writes themselves are ordered by grid
    // completion (the next kernel's griddepcontrol.wait / stream order), as in CUTLASS' tma_store_wait<0>()
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    if (bad && overflow) *overflow = 1;
    if (warp == 2 && lane == 0) trace_stamp(trace, 5);
  }
  tc_fence_before();
  cluster_sync_all();          // neither CTA may free TMEM or exit while the peer can still touch it
  if (warp == 0 && lane == 0) trace_stamp(trace, 6);
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
  static_assert(cfg::STAGE % 1024 == 0, "stages must keep the 1024-byte swizzle alignment");
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
  CUtensorMap tmA, tmW, tmO0, tmO1, tmO2;
  RB_TRY(tensor_map_2d(g.A, a_rows, g.K, BM, ELEM_BYTES, &tmA, 0));
  RB_TRY(tensor_map_2d(g.W, w_rows, g.K, BN / 2, ELEM_BYTES, &tmW, 0));
  // output maps: boxes of 32 rows x 128 bytes, clipped by TMA at row M / column N.
  //   EPI_STORE / EPI_RESIDUAL: O0 = C            EPI_RELU_ACT: O0, O1 = activation planes
  //   EPI_RESID_NORM: O0 = C, O1, O2 = activation planes
  const bool planes_only = g.epilogue == EPI_RELU_ACT || g.epilogue == EPI_PLANES;
  const bool planes_out = planes_only || g.epilogue == EPI_RESID_NORM;
  if (planes_out) {
    RB_REQUIRE((g.N * ELEM_BYTES) % 16 == 0, "N=%lld: activation rows must be multiples of 16 bytes", (long long)g.N);
    RB_REQUIRE(g.act.base != nullptr, "this epilogue needs an activation output buffer");
  }
  if (g.epilogue == EPI_RESID_NORM) {
    RB_REQUIRE(g.N % 64 == 0 && g.nf.ss_prev && g.nf.ss_out && g.nf.np == g.N / 64,
               "EPI_RESID_NORM needs N %% 64 == 0 and the NormFold tables (np = N / 64)");
  }
  if (!planes_only) RB_TRY(tensor_map_2d(g.C, g.M, g.N, 32, 4, &tmO0, g.ldc));
  if (planes_out) {
    CUtensorMap p0, p1;
    RB_TRY(tensor_map_2d(g.act.base, g.M, g.N, 32, ELEM_BYTES, &p0, g.N));
    p1 = p0;
    if (cfg::PLANES == 2)
      RB_TRY(tensor_map_2d(static_cast<const char*>(g.act.base) + g.act.plane * ELEM_BYTES, g.M, g.N, 32, ELEM_BYTES,
                           &p1, g.N));
    if (planes_only) { tmO0 = p0; tmO1 = p1; tmO2 = p1; }
    else { tmO1 = p0; tmO2 = p1; }
  } else {
    tmO1 = tmO0;
    tmO2 = tmO0;
  }
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
  RB_CUDA(launch_pdl(kern, grid, dim3(kThreads), cfg::SMEM, s, tmA, tmW, tmO0, tmO1, tmO2, (const float*)g.C, g.ldc,
                     g.nf, g.act.overflow, (int)prec_is_fp16(g.mode), (int)g.M, (int)g.N, (int)g.K, (int)a_plane_rows, (int)w_plane_rows,
                     g.epilogue, make_idesc_pair(fmt, BN), g.out_scale, num_n_tiles, num_tiles, g_gemm_trace));
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
  const int64_t tiles256 = m_pairs * ceil_div(g.N, 256);
  const int64_t cost256 = (int64_t)ceil_div(tiles256, sm_pairs) * 256;
  const int64_t cost128 = (int64_t)ceil_div(m_pairs * ceil_div(g.N, 128), sm_pairs) * 128;
  // Many rounds per SM pair (the forced tail: M = 71680): quantisation no longer matters and the 128-wide tiles are
  // bound by operand delivery from L2 (ncu: 9.2-9.7 TB/s L2->SM on every tail GEMM); 256-wide tiles move a third
  // fewer operand bytes per output. RB200_BN=128|256 forces a width.
  static const int force = []() {
    const char* e = getenv("RB200_BN");
    return e ? atoi(e) : 0;
  }();
  const bool wide = force ? force == 256 : (cost256 <= cost128 || tiles256 >= 4 * (int64_t)sm_pairs);
  // Small batches (few hundred rows: the reference's shipped --batch_size 1 / 4 launches): every tile has an SM pair
  // to itself and the launch lasts as long as ONE tile's K loop, which 128-wide tiles run at the shared-memory
  // operand bandwidth (0.55 us per 64-deep k-block). 64-wide tiles halve the MMA work per k-block; taken only while
  // all of them still fit one round.
  const bool narrow = force ? force == 64 : m_pairs * ceil_div(g.N, 64) <= sm_pairs;
  if (narrow) return launch_cfg2<ELEM_BYTES, NTERMS, 64>(g, s);
  if (g.N >= 256 && wide) return launch_cfg2<ELEM_BYTES, NTERMS, 256>(g, s);
  return launch_cfg2<ELEM_BYTES, NTERMS, 128>(g, s);
}

}  // namespace

int launch_gemm_sm100_2cta(const GemmArgs& g, cudaStream_t s) {
  const int elem = prec_elem_bytes(g.mode);
  RB_REQUIRE(g.K % (16 / elem) == 0, "K=%lld must be a multiple of %d for TMA", (long long)g.K, 16 / elem);
  RB_REQUIRE(g.N % 4 == 0, "N=%lld must be a multiple of 4", (long long)g.N);
  RB_REQUIRE(g.epilogue == EPI_RELU_ACT || g.epilogue == EPI_PLANES || g.ldc % 4 == 0, "ldc must be a multiple of 4");
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
