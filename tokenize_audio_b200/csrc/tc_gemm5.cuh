// tcgen05 implicit-GEMM, fifth generation: CTA PAIRS (cta_group::2). Two CTAs of a cluster (the two SMs of a TPC)
// compute two 128-row tiles against the SAME BNP weight columns with one M = 256 MMA: each CTA stages its own 128
// activation rows and only HALF of the weight k-block (BNP/2 rows); the tensor cores of both SMs read both halves.
//
// Why: the single-CTA kernel (tc_gemm2) needs 64 KB of operands per 128 x 128 x 32 k-block = 85 B/clk/SM at full tensor
// rate, and ncu shows the xbar -> SM stream pinned at ~12 TB/s chip-wide (~43 B/clk/SM, the L2 slice throughput cap)
// with the tensor pipe 50-65 % busy: the wide layers are L2 -> shared-memory bandwidth bound. A pair tile of
// 2 x 128 rows x 256 columns moves the same 64 KB per CTA for twice the MMA work (43 B/clk/SM), and reads each operand
// byte from shared memory once per 256 x 256 x 8 MMA instead of once per 128 x 128 x 8.
//
//   grid    = 2 x clusters, persistent; cluster c walks pair tiles pid = c, c + #clusters, ...
//   pair tile pid -> (n-tile, q): CTA rank r takes the 128-row tile t = 2q + r of the list (m-tile, item) -- the two
//             halves of the M = 256 MMA need not be adjacent rows (each CTA loads its own A rows), so items pair up
//             tile by tile and a short item wastes at most its own last partial tile.
//   warp 0  = TMA producer (both CTAs; all loads signal the LEADER's full barrier), warp 1 = MMA issuer (leader CTA
//             only: tcgen05.mma.cta_group::2, commits multicast to both CTAs), warps 2-3 idle, warps 4-19 = epilogue
//             (each CTA drains its own TMEM: lanes = its 128 rows). setmaxnreg: 32 / 112 registers -- setmaxnreg only
//             moves registers inside the CTA's launch allocation (640 x 96 = 128 x 32 + 512 x 112): asking for more
//             blocks the last warps in setmaxnreg.inc forever.
//   TMEM    = 2 (BNP = 256) or 4 chunk buffers x BNP columns. One accumulator per chunk: hi*hi, hi*lo and lo*hi all land in it and it is
//             drained into fp32 registers every K = 128 (round-to-nearest adds), as in tc_gemm2.
#pragma once
#include "tc_gemm2.cuh"

namespace mimi {
namespace tcp {

using tc::Epilogue;
using tc::kBK;
using tc::kBM;
using tc::kChunkKB;
using tc::kUmmaK;
using tc2::Sched;                                   // ntn = N / BNP here

constexpr int kEpiWarps = 16;
constexpr int kEW0 = 4;                             // first epilogue warp
constexpr int kThreads = 32 * (kEW0 + kEpiWarps);   // 640
constexpr int kSmemMax = 232448;

// Debug build only (-DMIMI_TCP_DEBUG): progress marks of cluster 0 in mapped host memory, readable while a kernel hangs
#ifdef MIMI_TCP_DEBUG
__device__ unsigned* g_tcp_marks = nullptr;
#define TCP_MARK(slot, v)                                                                   \
  do {                                                                                      \
    if (g_tcp_marks && cid == 0) {                                                          \
      ((volatile unsigned*)g_tcp_marks)[rank * 16 + (slot)] = (unsigned)(v);                \
      __threadfence_system();                                                               \
    }                                                                                       \
  } while (0)
#else
#define TCP_MARK(slot, v)
#endif

// LOB = 1 (mode 7, TF32 generation): the stage holds A_hi (fp32, SWIZZLE_128B, 16 KB), A_lo (bf16, SWIZZLE_64B rows of 64 B,
// 8 KB), W_hi, W_lo (fp32) and W_hib = bf16(W_hi) (SWIZZLE_64B); A_hi*W_hi and A_hi*W_lo run on kind::tf32 (four K = 8 steps per
// k-block each), A_lo*W_hib on kind::f16 (two K = 16 steps): 5 bf16-rate pass units per MAC, 6 bytes per activation element.
// LOB = 3 (mode 9, the fp16 generation, default): every operand is an fp16 array in SWIZZLE_64B tiles -- A_hi16, A_lo16
// (= (x - hi) * 2048), W_hi16, W_lo16 and W_hs16 = W_hi16 / 2048 of the row-scaled weights (common.cuh: split_f16) -- and all
// three products A_hi*W_hi + A_hi*W_lo + A_lo*W_hs run on kind::f16: 6 MMAs of K = 16 per k-block (3 tensor passes at the
// 16-bit rate), and the stage shrinks from 64 KB to 40 KB.
template <int BNP, int LOB>
struct Cfg {
  static_assert(LOB == 1 || LOB == 3, "LOB");
  static constexpr int WB = BNP / 2;                                 // weight rows staged by each CTA
  static constexpr int A_BYTES = kBM * kBK * 4;                      // 16 KB (fp32 tile)
  static constexpr int W_BYTES = WB * kBK * 4;
  // LOB 1: [A_hi | A_lob | W_hi | W_lo | W_hib]      LOB 3: [A_hi16 | A_lo16 | W_hi16 | W_lo16 | W_hs16]
  static constexpr int OFF_ALO = LOB == 3 ? A_BYTES / 2 : A_BYTES;
  static constexpr int OFF_WHI = OFF_ALO + A_BYTES / 2;
  static constexpr int OFF_WLO = OFF_WHI + (LOB == 3 ? W_BYTES / 2 : W_BYTES);          // fp32 W_lo (LOB 1), W_lo16 (LOB 3)
  static constexpr int OFF_W3 = OFF_WLO + (LOB == 3 ? W_BYTES / 2 : W_BYTES);           // bf16(W_hi) (LOB 1), W_hs16 (LOB 3)
  static constexpr int STAGE = OFF_W3 + W_BYTES / 2;                                    // per CTA
  static constexpr int PC = 16;
  static constexpr int STG_WARP = 32 * PC * 4;
  static constexpr int STG = kEpiWarps * STG_WARP;
  static constexpr int BAR_BYTES = 512;
  static constexpr int STAGES_RAW = (kSmemMax - 1024 - STG - BAR_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = 1024 + STAGES * STAGE + STG + BAR_BYTES;
  // accumulator chunk buffers in TMEM: two for 256-column tiles (all 512 columns); four for the narrower tiles, so that the MMAs
  // of the next tile (up to 4 chunks = K 512) run while the epilogue warps are busy finishing the previous one
  static constexpr int NBUF = BNP == 256 ? 2 : 4;
  static constexpr int TMEM_COLS = NBUF * BNP;
  static constexpr int HALF = BNP / (kEpiWarps / 4);                 // accumulator columns per epilogue thread
  static_assert(BNP == 64 || BNP == 128 || BNP == 256, "BNP");
  static_assert(STAGES >= 3, "ring too shallow");
  static_assert(STAGE % 1024 == 0 && OFF_WHI % 1024 == 0 && OFF_WLO % 512 == 0 && OFF_W3 % 512 == 0 && OFF_ALO % 512 == 0,
                "operand alignment");
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of the same shared-memory location in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  __syncwarp();                                      // .aligned: the whole warp must execute the barrier together
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// The same without release semantics, for "this TMEM buffer is drained": nothing in generic memory has to become visible
// to the waiter (the tcgen05 loads are ordered by tcgen05.fence::before_thread_sync), and a cluster-scope release makes
// the arriving lane wait for every global store its warp still has in flight from the previous tile's finish -- 17 % of
// the warp samples of D1 sat in that fence (ERRBAR + the arrive, profiles/r01_final3_ncu.md).
__device__ __forceinline__ void mbar_arrive_cluster_relaxed(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: destination in the executing CTA, completion bytes on a barrier that may live in the peer
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
// M = 256 across the pair; operands at the same shared-memory offsets in both CTAs (A: own 128 rows, B: own N/2 rows)
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(tc::kDescHi)
      : "memory");
}
// the same for bf16 operands in SWIZZLE_64B K-major tiles (rows of 64 B, 8-row groups 512 B apart): kind::f16, K = 16
constexpr uint32_t kDescHi64 = (uint32_t)(512 >> 4) | (1u << 14) | (4u << 29);
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// kind::f16 with fp16 A and B (format 0), fp32 accumulator
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(kDescHi64)
      : "memory");
}
// arrive (once every MMA issued so far has completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"(mask) : "memory");
}

template <int BNP, int LOB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tcp_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                const __grid_constant__ CUtensorMap tmW_3, int K, const Epilogue ep, const Sched sc) {
  using C = Cfg<BNP, LOB>;
  constexpr int STAGES = C::STAGES;
  constexpr int STAGE = C::STAGE;
  constexpr int HALF = C::HALF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + STAGES * STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + C::STG);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;                                // one per CTA (multicast commit)
  constexpr uint32_t NBUF = C::NBUF;
  uint64_t* acc_full = empty_bar + STAGES;                                // [NBUF] one per CTA (multicast commit)
  uint64_t* acc_empty = acc_full + NBUF;                                  // [NBUF] leader only: both CTAs' epilogue warps arrive
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(acc_empty + NBUF);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nkb = K / kBK;
  const int ckb = ep.chunk_kb > 0 ? ep.chunk_kb : kChunkKB;
  const int nchunks = (nkb + ckb - 1) / ckb;
  const int ntiles = tc2::sched_tiles(sc);           // 128-row tiles (m-tile major, item minor; compact list if given)
  const int npairs = ((ntiles + 1) >> 1) * sc.ntn;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi); tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmW_hi); tc::prefetch_tmap(&tmW_lo);
    tc::prefetch_tmap(&tmW_3);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (uint32_t s = 0; s < NBUF; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], 2 * kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // both CTAs of the pair run the allocation (same warp index, same destination offset)
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();                                    // barriers of both CTAs initialised before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;
  if (threadIdx.x == 0) { TCP_MARK(0, 1); TCP_MARK(8, tmem_base); }

  // tile r (0 / 1) of pair tile pid -> (item, first row, first column); false when the tile does not exist
  auto decode = [&](int pid, int r, int& b, int& m0, int& n0, int& Lout) {
    const int nt = pid % sc.ntn;
    const int t = 2 * (pid / sc.ntn) + r;
    n0 = nt * BNP;
    b = 0; m0 = 0; Lout = 0;
    if (t >= ntiles) return false;
    return tc2::sched_tile(sc, ep, t, b, m0, Lout);
  };
  // this CTA's tile of the pair, and whether the pair has any work (identical decision in every role of both CTAs)
  auto decode_pair = [&](int pid, int& b, int& m0, int& n0, int& Lout, bool& mine) {
    int b1, m1, n1, L1;
    const bool v0 = decode(pid, 0, b, m0, n0, Lout);
    const bool v1 = decode(pid, 1, b1, m1, n1, L1);
    mine = v0;
    if (rank) { b = b1; m0 = m1; Lout = L1; mine = v1; }
    return v0 || v1;
  };

  if (warp < kEW0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      if (tc::elect_one()) {
        const uint32_t full_leader = mapa(tc::smem_u32(full_bar), 0);
        const uint32_t smem_u = tc::smem_u32(smem);
        // The producer is ONE thread; every instruction between two TMA issues is on the critical path of the narrow layers
        // (a k-block of a 128-column pair tile is only 384 clocks of MMA work). Ring position and k-block order are therefore
        // walked incrementally: no division or modulo per k-block (the first version spent ~200 instructions per k-block here and
        // starved the tensor pipe of D1 / R2a: epilogue warps waited 15 % of their samples for accumulators).
        uint32_t s = 0, ring_phase = 0;
        for (int pid = cid; pid < npairs; pid += ncl) {
          int b, m0, n0, Lout;
          bool mine;
          if (!decode_pair(pid, b, m0, n0, Lout, mine)) continue;
          const int wrow = n0 + (int)rank * C::WB;
          int dq = 0, cb = 0, ph = 0;                      // tap group, channel panel, tap phase of the next k-block (Sched::G > 1)
          for (int kb = 0; kb < nkb; ++kb) {
            int kx = kb * kBK;
            if (sc.G > 1) {                                // i-th k-block visited = (ph + s * dq) * cp + cb, dq fastest (tc2::kblock_order)
              kx = ((ph + sc.s * dq) * sc.cp + cb) * kBK;
              if (++dq == sc.G) { dq = 0; if (++cb == sc.cp) { cb = 0; ++ph; } }
            }
            tc::mbar_wait(&empty_bar[s], ring_phase ^ 1u);
            if (rank == 0) tc::mbar_expect_tx(&full_bar[s], 2 * STAGE);
            const uint32_t st = smem_u + s * STAGE;
            const uint32_t fb = full_leader + 8u * s;
            tma_load_3d_pair(st, &tmA_hi, fb, kx, m0, b);
            tma_load_3d_pair(st + C::OFF_ALO, &tmA_lo, fb, kx, m0, b);
            tma_load_2d_pair(st + C::OFF_WHI, &tmW_hi, fb, kx, wrow);
            tma_load_2d_pair(st + C::OFF_WLO, &tmW_lo, fb, kx, wrow);
            tma_load_2d_pair(st + C::OFF_W3, &tmW_3, fb, kx, wrow);
            if (++s == STAGES) { s = 0; ring_phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      constexpr uint32_t idesc = tc::make_idesc(2 * kBM, BNP);
      const uint32_t smem_base_u32 = tc::smem_u32(smem);
      uint32_t s = 0, ring_phase = 0, cc = 0;
      for (int pid = cid; pid < npairs; pid += ncl) {
        int b, m0, n0, Lout;
        bool mine;
        if (!decode_pair(pid, b, m0, n0, Lout, mine)) continue;
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const uint32_t buf = cc % NBUF;
          if (lane == 0) TCP_MARK(3, cc | 0x80000000u);
          tc::mbar_wait(&acc_empty[buf], ((cc / NBUF) & 1u) ^ 1u);       // drained (by both CTAs) NBUF chunks ago
          if (lane == 0) TCP_MARK(3, cc + 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_acc = tmem_base + buf * BNP;
          const int kb_end = min(nkb, (c + 1) * ckb);
          for (int kb = c * ckb; kb < kb_end; ++kb) {
            tc::mbar_wait(&full_bar[s], ring_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_ahi = tc::desc_lo(smem_base_u32 + s * STAGE);
            constexpr uint32_t kAlo = C::OFF_ALO >> 4, kWhi = C::OFF_WHI >> 4, kWlo = C::OFF_WLO >> 4, kW3 = C::OFF_W3 >> 4;
            const bool first_in_chunk = kb == c * ckb;
            if (tc::elect_one()) {
              if (LOB == 3) {
                constexpr uint32_t idesc_h = make_idesc_f16(2 * kBM, BNP);
#pragma unroll
                for (int k = 0; k < 2; ++k) {   // 32 fp16 = two K = 16 steps of 32 B
                  umma_bf16_pair(tmem_acc, d_ahi + 2 * k, d_ahi + kWhi + 2 * k, idesc_h, !(first_in_chunk && k == 0));
                  umma_bf16_pair(tmem_acc, d_ahi + 2 * k, d_ahi + kWlo + 2 * k, idesc_h, 1u);
                  umma_bf16_pair(tmem_acc, d_ahi + kAlo + 2 * k, d_ahi + kW3 + 2 * k, idesc_h, 1u);
                }
              } else {
#pragma unroll
                for (int k = 0; k < kBK / kUmmaK; ++k) {
                  umma_tf32_pair(tmem_acc, d_ahi + 2 * k, d_ahi + kWhi + 2 * k, idesc, !(first_in_chunk && k == 0));
                  umma_tf32_pair(tmem_acc, d_ahi + 2 * k, d_ahi + kWlo + 2 * k, idesc, 1u);
                }
                constexpr uint32_t idesc_b = make_idesc_bf16(2 * kBM, BNP);
#pragma unroll
                for (int k = 0; k < 2; ++k)     // 32 bf16 = two K = 16 steps of 32 B
                  umma_bf16_pair(tmem_acc, d_ahi + kAlo + 2 * k, d_ahi + kW3 + 2 * k, idesc_b, 1u);
              }
              umma_commit_pair(&empty_bar[s]);
              if (kb + 1 == kb_end) umma_commit_pair(&acc_full[buf]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ring_phase ^= 1u; }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ---- epilogue warps: TMEM lane quarter = warp % 4, column slice = (warp - kEW0) / 4 ---------------------------------
    const int ew = warp - kEW0;
    const int quarter = warp & 3;
    const int col0 = (ew >> 2) * HALF;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t stg = tc::smem_u32(stg_base + ew * C::STG_WARP);
    const uint32_t acc_empty_leader = mapa(tc::smem_u32(acc_empty), 0);
    uint32_t cc = 0;
    for (int pid = cid; pid < npairs; pid += ncl) {
      int b, m0, n0, Lout;
      bool mine;
      if (!decode_pair(pid, b, m0, n0, Lout, mine)) continue;
      float acc[HALF];
#pragma unroll
      for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
      if (mine) tc2::prefetch_residual<HALF, C::PC>(ep, b, m0 + quarter * 32, n0 + col0, Lout, lane);
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const uint32_t buf = cc % NBUF;
        if (threadIdx.x == 32 * kEW0) TCP_MARK(4, cc | 0x80000000u);
        tc::mbar_wait(&acc_full[buf], (cc / NBUF) & 1u);
        if (threadIdx.x == 32 * kEW0) TCP_MARK(4, cc + 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc2::drain_add<HALF>(tmem_base + lane_off + buf * BNP + (uint32_t)col0, acc);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(acc_empty_leader + 8u * buf);
      }
      if (mine) tc2::finish_tile<HALF, C::PC, LOB>(ep, acc, b, m0 + quarter * 32, n0 + col0, Lout, stg, lane);
      if (threadIdx.x == 32 * kEW0) TCP_MARK(5, pid + 1);
    }
  }
  if (lane == 0) TCP_MARK(9 + (warp < 5 ? warp : 5), 2);
  // nobody leaves (or frees TMEM) while the peer may still read this CTA's shared memory or signal its barriers
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync();
  if (threadIdx.x == 0) TCP_MARK(0, 3);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
  }
}

}  // namespace tcp
}  // namespace mimi
