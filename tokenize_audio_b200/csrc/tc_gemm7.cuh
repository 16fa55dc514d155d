// The CTA-pair tcgen05 GEMM of tc_gemm5.cuh for convolutions whose taps share input rows (fp16 generation only).
//
// A conv with k = G * s taps (G = 2: the strided convs, k = 2 s; G = 3: the k = 3, s = 1 convs of the residual blocks) reads,
// for tap tau + s, exactly the rows tap tau reads for the NEXT output row: in the im2col view of tc_gemm5 the k-blocks
// (tau, channel panel c) and (tau + s, c) are the same 128 x 32 tile shifted by one row, and the kernel there fetched it twice
// (three times for k = 3) -- the second fetch is an L2 hit, but ncu shows these layers bound by the L2 -> shared-memory stream
// (xbar 9-10 of ~12 TB/s, lts 60 %, tensor pipe 58-65 % on D2 / R3a / R4a; D1 at 7.4 TB/s on top of its DRAM traffic).
// Here a pipeline stage holds ONE activation tile of 128 + G - 1 rows per (tap phase, channel panel) and the G weight
// k-blocks that go with it; tap group member g runs its MMAs through an A descriptor whose start address is g rows (g x 64
// bytes) further -- the row-shift property front_fused.cuh relies on, in its SWIZZLE_64B form. Activation bytes through L2 ->
// shared memory drop by G, the stage count per tile by G, everything else (pair tiles, multicast commits, chunked accumulation,
// tile finish) is tc_gemm5's.
//
// The activation maps of these launches have one extra box row per extra tap (rows m0 .. m0 + 127 + G - 1) and G - 1 extra
// rows in their row dimension: the row past an item's last output row holds real input rows (the first taps of it are read by
// the last output row through the shifted descriptor), which a map ending at the last output row would zero-fill.
#pragma once
#include "tc_gemm5.cuh"

namespace mimi {
namespace tcg {

using tc::Epilogue;
using tc::kBK;
using tc::kBM;
using tc2::Sched;                                   // G = taps per group, s = conv stride, cp = C_in / 32 channel panels
using tcp::kEpiWarps;
using tcp::kEW0;
using tcp::kSmemMax;
using tcp::kThreads;

template <int BNP, int TG>
struct Cfg {
  static_assert(TG == 2 || TG == 3, "tap group");
  static_assert(BNP == 64 || BNP == 128, "BNP");
  static constexpr int WB = BNP / 2;                                 // weight rows staged by each CTA
  static constexpr int A_ROWS = kBM + TG - 1;
  static constexpr int A_TILE = 9216;                                // A_ROWS x 64 B (<= 8320) rounded up to a multiple of 1024
  static constexpr int W_TILE = WB * kBK * 2;
  static constexpr int OFF_ALO = A_TILE;
  static constexpr int OFF_W = 2 * A_TILE;                           // then per tap g: [W_hi | W_lo | W_hs]
  static constexpr int STAGE = OFF_W + TG * 3 * W_TILE;              // per CTA
  static constexpr int TX = 2 * A_ROWS * kBK * 2 + TG * 3 * W_TILE;  // bytes one CTA's loads of a stage deliver
  static constexpr int PC = 16;
  static constexpr int STG_WARP = 32 * PC * 4;
  static constexpr int STG = kEpiWarps * STG_WARP;
  static constexpr int BAR_BYTES = 512;
  static constexpr int STAGES_RAW = (kSmemMax - 1024 - STG - BAR_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = 1024 + STAGES * STAGE + STG + BAR_BYTES;
  static constexpr int NBUF = 4;                                     // accumulator chunk buffers in TMEM
  static constexpr int TMEM_COLS = NBUF * BNP;
  static constexpr int HALF = BNP / (kEpiWarps / 4);
  static constexpr int CST = 2;                                      // stages per accumulation chunk: K = 128 (TG 2) / 192 (TG 3)
  static_assert(STAGES >= 3, "ring too shallow");
  static_assert(STAGE % 1024 == 0 && W_TILE % 512 == 0 && A_ROWS * kBK * 2 <= A_TILE, "operand alignment");
};

template <int BNP, int TG>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tcp_taps_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                const __grid_constant__ CUtensorMap tmW_3, const Epilogue ep, const Sched sc) {
  using C = Cfg<BNP, TG>;
  constexpr int STAGES = C::STAGES;
  constexpr int STAGE = C::STAGE;
  constexpr int HALF = C::HALF;
  constexpr int LOB = 3;
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
  const uint32_t rank = tcp::cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nst = sc.s * sc.cp;                      // stages per tile: (tap phase, channel panel); K = TG * nst * 32
  const int nchunks = (nst + C::CST - 1) / C::CST;
  const int ntiles = tc2::sched_tiles(sc);
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
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tcp::cluster_sync();                               // barriers of both CTAs initialised before any remote arrive
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;

  auto decode = [&](int pid, int r, int& b, int& m0, int& n0, int& Lout) {
    const int nt = pid % sc.ntn;
    const int t = 2 * (pid / sc.ntn) + r;
    n0 = nt * BNP;
    b = 0; m0 = 0; Lout = 0;
    if (t >= ntiles) return false;
    return tc2::sched_tile(sc, ep, t, b, m0, Lout);
  };
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
        const uint32_t full_leader = tcp::mapa(tc::smem_u32(full_bar), 0);
        const uint32_t smem_u = tc::smem_u32(smem);
        const int tap_step = sc.s * sc.cp * kBK;           // inner-coordinate distance of tap tau + s from tap tau
        uint32_t s = 0, ring_phase = 0;
        for (int pid = cid; pid < npairs; pid += ncl) {
          int b, m0, n0, Lout;
          bool mine;
          if (!decode_pair(pid, b, m0, n0, Lout, mine)) continue;
          const int wrow = n0 + (int)rank * C::WB;
          for (int i = 0, kx = 0; i < nst; ++i, kx += kBK) {     // kx = (ph * cp + cb) * 32: the k-block of the group's first tap
            tc::mbar_wait(&empty_bar[s], ring_phase ^ 1u);
            if (rank == 0) tc::mbar_expect_tx(&full_bar[s], 2 * C::TX);
            const uint32_t st = smem_u + s * STAGE;
            const uint32_t fb = full_leader + 8u * s;
            tcp::tma_load_3d_pair(st, &tmA_hi, fb, kx, m0, b);
            tcp::tma_load_3d_pair(st + C::OFF_ALO, &tmA_lo, fb, kx, m0, b);
#pragma unroll
            for (int g = 0; g < TG; ++g) {
              const uint32_t wd = st + C::OFF_W + g * 3 * C::W_TILE;
              tcp::tma_load_2d_pair(wd, &tmW_hi, fb, kx + g * tap_step, wrow);
              tcp::tma_load_2d_pair(wd + C::W_TILE, &tmW_lo, fb, kx + g * tap_step, wrow);
              tcp::tma_load_2d_pair(wd + 2 * C::W_TILE, &tmW_3, fb, kx + g * tap_step, wrow);
            }
            if (++s == STAGES) { s = 0; ring_phase ^= 1u; }
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      constexpr uint32_t idesc_h = tcp::make_idesc_f16(2 * kBM, BNP);
      const uint32_t smem_base_u32 = tc::smem_u32(smem);
      uint32_t s = 0, ring_phase = 0, cc = 0;
      for (int pid = cid; pid < npairs; pid += ncl) {
        int b, m0, n0, Lout;
        bool mine;
        if (!decode_pair(pid, b, m0, n0, Lout, mine)) continue;
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const uint32_t buf = cc % NBUF;
          tc::mbar_wait(&acc_empty[buf], ((cc / NBUF) & 1u) ^ 1u);       // drained (by both CTAs) NBUF chunks ago
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_acc = tmem_base + buf * BNP;
          const int i_end = min(nst, (c + 1) * C::CST);
          for (int i = c * C::CST; i < i_end; ++i) {
            tc::mbar_wait(&full_bar[s], ring_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_ahi = tc::desc_lo(smem_base_u32 + s * STAGE);
            const bool first_in_chunk = i == c * C::CST;
            if (tc::elect_one()) {
#pragma unroll
              for (int g = 0; g < TG; ++g) {
                // tap group member g: the same staged rows, g rows (g x 64 bytes) further
                const uint32_t a_hi = d_ahi + (uint32_t)(g * 4), a_lo = a_hi + (C::OFF_ALO >> 4);
                const uint32_t w = d_ahi + ((C::OFF_W + g * 3 * C::W_TILE) >> 4);
#pragma unroll
                for (int k = 0; k < 2; ++k) {   // 32 fp16 = two K = 16 steps of 32 B
                  tcp::umma_bf16_pair(tmem_acc, a_hi + 2 * k, w + 2 * k, idesc_h, !(first_in_chunk && g == 0 && k == 0));
                  tcp::umma_bf16_pair(tmem_acc, a_hi + 2 * k, w + (C::W_TILE >> 4) + 2 * k, idesc_h, 1u);
                  tcp::umma_bf16_pair(tmem_acc, a_lo + 2 * k, w + ((2 * C::W_TILE) >> 4) + 2 * k, idesc_h, 1u);
                }
              }
              tcp::umma_commit_pair(&empty_bar[s]);
              if (i + 1 == i_end) tcp::umma_commit_pair(&acc_full[buf]);
            }
            __syncwarp();
            if (++s == STAGES) { s = 0; ring_phase ^= 1u; }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    // ---- epilogue warps: TMEM lane quarter = warp % 4, column slice = (warp - kEW0) / 4 (as in tc_gemm5) ----------------
    const int ew = warp - kEW0;
    const int quarter = warp & 3;
    const int col0 = (ew >> 2) * HALF;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t stg = tc::smem_u32(stg_base + ew * C::STG_WARP);
    const uint32_t acc_empty_leader = tcp::mapa(tc::smem_u32(acc_empty), 0);
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
        tc::mbar_wait(&acc_full[buf], (cc / NBUF) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc2::drain_add<HALF>(tmem_base + lane_off + buf * BNP + (uint32_t)col0, acc);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tcp::mbar_arrive_cluster_relaxed(acc_empty_leader + 8u * buf);
      }
      if (mine) tc2::finish_tile<HALF, C::PC, LOB>(ep, acc, b, m0 + quarter * 32, n0 + col0, Lout, stg, lane);
    }
  }
  // nobody leaves (or frees TMEM) while the peer may still read this CTA's shared memory or signal its barriers
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tcp::cluster_sync();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
  }
}

}  // namespace tcg
}  // namespace mimi
