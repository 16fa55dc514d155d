// Fused level-1 residual block (MimiResnetBlock, modeling_mimi.py:412-451) on the CTA-pair GEMM of tc_gemm5.cuh, bf16-lo
// generation:   h2 = split(ELU(d1 + b_b + W_b * ELU(b_a + W_a (*) ELU(d1))))
//   conv a: 128 -> 64 channels, k = 3 (K = 384) over the ELU'd hi / lo(bf16) split of d1 (written by D1);
//   conv b: 64 -> 128 channels, k = 1 (K = 64), + skip (raw d1), then ELU + split for the next strided conv.
// Unfused, the 64-channel intermediate crosses HBM twice (6 bytes per element, 6000 rows per audio-second) and conv b is a
// launch of its own that moves 7 GB per bench step for 0.1 ms of tensor work. Here a pair tile (2 x 128 rows) runs conv a
// through the k-block ring as usual; the epilogue warps turn the accumulators into the K-major operand of conv b in shared
// memory (fp32 hi panels in SWIZZLE_128B, bf16 lo panels in SWIZZLE_64B), the MMA warp runs conv b against weights that
// stay resident in shared memory, and the usual tile finish (bias, skip, ELU, split, coalesced stores) follows.
#pragma once
#include "tc_gemm5.cuh"

namespace mimi {
namespace tcr {

using tc::Epilogue;
using tc::kBK;
using tc::kBM;
using tc2::Sched;
using tcp::kEpiWarps;
using tcp::kEW0;
using tcp::kThreads;

constexpr int kNA = 64;                              // conv a output channels = conv b K
constexpr int kNB = 128;                             // conv b output channels
using CA = tcp::Cfg<kNA, 1>;                         // ring stage of conv a: A_hi | A_lob | W_hi | W_lo | W_hib
constexpr int kStages = 3;
constexpr int kRHi = 2 * kBM * 128;                  // conv b operand, hi: 2 panels (32 channels each) x 128 rows x 128 B
constexpr int kRLo = 2 * kBM * 64;                   //                 lo (bf16): 2 panels x 128 rows x 64 B
constexpr int kWbHi = 2 * (kNB / 2) * 128;           // resident conv b weights of this CTA (64 rows): hi, 2 panels
constexpr int kWbLo = kWbHi;
constexpr int kWbHb = kWbHi / 2;
constexpr int kStg = kEpiWarps * 32 * 16 * 4;        // 32 KB transpose staging of the tile finish
constexpr int OFF_R = kStages * CA::STAGE;
constexpr int OFF_WB = OFF_R + kRHi + kRLo;
constexpr int OFF_STG = OFF_WB + kWbHi + kWbLo + kWbHb;
constexpr int OFF_BAR = OFF_STG + kStg;
constexpr int kSmem = 1024 + OFF_BAR + 512;
constexpr int kTmemCols = 256;                       // conv a: 2 chunk buffers x 64 columns; conv b: 128 columns at +128
static_assert(kSmem <= tcp::kSmemMax, "shared memory");
static_assert(OFF_R % 1024 == 0 && OFF_WB % 1024 == 0, "operand alignment");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tcr_resblock_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmWa_hi, const __grid_constant__ CUtensorMap tmWa_lo,
                    const __grid_constant__ CUtensorMap tmWa_hib, const __grid_constant__ CUtensorMap tmWb_hi,
                    const __grid_constant__ CUtensorMap tmWb_lo, const __grid_constant__ CUtensorMap tmWb_hib, int Ka,
                    const float* __restrict__ bias_a, const Epilogue ep, const Sched sc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);       // leader only
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* acc_full = empty_bar + kStages;                               // [2] conv a chunk ready (multicast)
  uint64_t* acc_empty = acc_full + 2;                                     // [2] leader: conv a chunk drained by both CTAs
  uint64_t* r_ready = acc_empty + 2;                                      // leader: conv b operand written by both CTAs
  uint64_t* b_full = r_ready + 1;                                         // conv b accumulator ready (multicast)
  uint64_t* wb_full = b_full + 1;                                         // leader: resident conv b weights landed
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(wb_full + 1);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  const uint32_t rank = tcp::cluster_ctarank();
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;
  const int nkb = Ka / kBK;
  const int ckb = tc::kChunkKB;
  const int nchunks = (nkb + ckb - 1) / ckb;
  const int ntiles = tc2::sched_tiles(sc);
  const int npairs = (ntiles + 1) >> 1;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi); tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmWa_hi); tc::prefetch_tmap(&tmWa_lo);
    tc::prefetch_tmap(&tmWa_hib); tc::prefetch_tmap(&tmWb_hi); tc::prefetch_tmap(&tmWb_lo); tc::prefetch_tmap(&tmWb_hib);
    for (int s = 0; s < kStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 2 * kEpiWarps); }
    tc::mbar_init(r_ready, 2 * kEpiWarps);
    tc::mbar_init(b_full, 1);
    tc::mbar_init(wb_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tcp::cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;

  auto decode_pair = [&](int pid, int& b, int& m0, int& Lout, bool& mine) {
    int b0 = 0, m00 = 0, L0 = 0, b1 = 0, m1 = 0, L1 = 0;
    const bool v0 = 2 * pid < ntiles && tc2::sched_tile(sc, ep, 2 * pid, b0, m00, L0);
    const bool v1 = 2 * pid + 1 < ntiles && tc2::sched_tile(sc, ep, 2 * pid + 1, b1, m1, L1);
    b = rank ? b1 : b0; m0 = rank ? m1 : m00; Lout = rank ? L1 : L0; mine = rank ? v1 : v0;
    return v0 || v1;
  };

  const uint32_t smem_u = tc::smem_u32(smem);
  if (warp < kEW0) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if (warp == 0) {
      if (tc::elect_one()) {
        const uint32_t full_leader = tcp::mapa(tc::smem_u32(full_bar), 0);
        // resident conv b weights: this CTA's 64 output channels, K = 64 as two panels
        {
          const uint32_t wbl = tcp::mapa(tc::smem_u32(wb_full), 0);
          if (rank == 0) tc::mbar_expect_tx(wb_full, 2 * (kWbHi + kWbLo + kWbHb));
          const int wrow = (int)rank * (kNB / 2);
          for (int pn = 0; pn < 2; ++pn) {
            tcp::tma_load_2d_pair(smem_u + OFF_WB + pn * (kWbHi / 2), &tmWb_hi, wbl, pn * kBK, wrow);
            tcp::tma_load_2d_pair(smem_u + OFF_WB + kWbHi + pn * (kWbLo / 2), &tmWb_lo, wbl, pn * kBK, wrow);
            tcp::tma_load_2d_pair(smem_u + OFF_WB + kWbHi + kWbLo + pn * (kWbHb / 2), &tmWb_hib, wbl, pn * kBK, wrow);
          }
        }
        uint32_t kbc = 0;
        for (int pid = cid; pid < npairs; pid += ncl) {
          int b, m0, Lout;
          bool mine;
          if (!decode_pair(pid, b, m0, Lout, mine)) continue;
          const int wrow = (int)rank * CA::WB;
          for (int kb = 0; kb < nkb; ++kb, ++kbc) {
            const uint32_t s = kbc % kStages;
            tc::mbar_wait(&empty_bar[s], ((kbc / kStages) & 1u) ^ 1u);
            if (rank == 0) tc::mbar_expect_tx(&full_bar[s], 2 * CA::STAGE);
            const uint32_t st = smem_u + s * CA::STAGE;
            const uint32_t fb = full_leader + 8u * s;
            const int kx = tc2::kblock_order(sc, kb) * kBK;
            tcp::tma_load_3d_pair(st, &tmA_hi, fb, kx, m0, b);
            tcp::tma_load_3d_pair(st + CA::OFF_ALO, &tmA_lo, fb, kx, m0, b);
            tcp::tma_load_2d_pair(st + CA::OFF_WHI, &tmWa_hi, fb, kx, wrow);
            tcp::tma_load_2d_pair(st + CA::OFF_WLO, &tmWa_lo, fb, kx, wrow);
            tcp::tma_load_2d_pair(st + CA::OFF_WHB, &tmWa_hib, fb, kx, wrow);
          }
        }
      }
    } else if (warp == 1 && rank == 0) {
      constexpr uint32_t idesc_a = tc::make_idesc(2 * kBM, kNA), idesc_ab = tcp::make_idesc_bf16(2 * kBM, kNA);
      constexpr uint32_t idesc_b = tc::make_idesc(2 * kBM, kNB), idesc_bb = tcp::make_idesc_bf16(2 * kBM, kNB);
      uint32_t kbc = 0, cc = 0, tc_ = 0;
      tc::mbar_wait(wb_full, 0);
      for (int pid = cid; pid < npairs; pid += ncl) {
        int b, m0, Lout;
        bool mine;
        if (!decode_pair(pid, b, m0, Lout, mine)) continue;
        // ---- conv a through the ring -------------------------------------------------------------------------------------
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const uint32_t buf = cc & 1u;
          tc::mbar_wait(&acc_empty[buf], ((cc >> 1) & 1u) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_acc = tmem_base + buf * kNA;
          const int kb_end = min(nkb, (c + 1) * ckb);
          for (int kb = c * ckb; kb < kb_end; ++kb, ++kbc) {
            const uint32_t s = kbc % kStages;
            tc::mbar_wait(&full_bar[s], (kbc / kStages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_ahi = tc::desc_lo(smem_u + s * CA::STAGE);
            constexpr uint32_t kAlo = CA::OFF_ALO >> 4, kWhi = CA::OFF_WHI >> 4, kWlo = CA::OFF_WLO >> 4, kWhb = CA::OFF_WHB >> 4;
            const bool first_in_chunk = kb == c * ckb;
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                tcp::umma_tf32_pair(tmem_acc, d_ahi + 2 * k, d_ahi + kWhi + 2 * k, idesc_a, !(first_in_chunk && k == 0));
                tcp::umma_tf32_pair(tmem_acc, d_ahi + 2 * k, d_ahi + kWlo + 2 * k, idesc_a, 1u);
              }
#pragma unroll
              for (int k = 0; k < 2; ++k) tcp::umma_bf16_pair(tmem_acc, d_ahi + kAlo + 2 * k, d_ahi + kWhb + 2 * k, idesc_ab, 1u);
              tcp::umma_commit_pair(&empty_bar[s]);
              if (kb + 1 == kb_end) tcp::umma_commit_pair(&acc_full[buf]);
            }
            __syncwarp();
          }
        }
        // ---- conv b: operand written by the epilogue warps of both CTAs, weights resident -------------------------------
        tc::mbar_wait(r_ready, tc_ & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tc::elect_one()) {
          const uint32_t d_r = tc::desc_lo(smem_u + OFF_R), d_w = tc::desc_lo(smem_u + OFF_WB);
          const uint32_t tmem_b = tmem_base + 128u;
#pragma unroll
          for (int pn = 0; pn < 2; ++pn) {
            const uint32_t a_hi = d_r + ((pn * (kRHi / 2)) >> 4), a_lo = d_r + ((kRHi + pn * (kRLo / 2)) >> 4);
            const uint32_t w_hi = d_w + ((pn * (kWbHi / 2)) >> 4), w_lo = d_w + ((kWbHi + pn * (kWbLo / 2)) >> 4);
            const uint32_t w_hb = d_w + ((kWbHi + kWbLo + pn * (kWbHb / 2)) >> 4);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              tcp::umma_tf32_pair(tmem_b, a_hi + 2 * k, w_hi + 2 * k, idesc_b, (uint32_t)((pn | k) != 0));
              tcp::umma_tf32_pair(tmem_b, a_hi + 2 * k, w_lo + 2 * k, idesc_b, 1u);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) tcp::umma_bf16_pair(tmem_b, a_lo + 2 * k, w_hb + 2 * k, idesc_bb, 1u);
          }
          tcp::umma_commit_pair(b_full);
        }
        __syncwarp();
        ++tc_;
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int ew = warp - kEW0;
    const int quarter = warp & 3;
    const int cs = ew >> 2;                                             // column slice: 16 of conv a's 64, 32 of conv b's 128
    const int row = quarter * 32 + lane;                                // tile row of this thread
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t stg = smem_u + OFF_STG + (uint32_t)ew * (32 * 16 * 4);
    const uint32_t acc_empty_leader = tcp::mapa(tc::smem_u32(acc_empty), 0);
    const uint32_t r_ready_leader = tcp::mapa(tc::smem_u32(r_ready), 0);
    float ba[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) ba[i] = __ldg(bias_a + cs * 16 + i);
    uint32_t cc = 0, tc_ = 0;
    for (int pid = cid; pid < npairs; pid += ncl) {
      int b, m0, Lout;
      bool mine;
      if (!decode_pair(pid, b, m0, Lout, mine)) continue;
      if (mine) tc2::prefetch_residual<32, 16>(ep, b, m0 + quarter * 32, cs * 32, Lout, lane);
      float acc[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const uint32_t buf = cc & 1u;
        tc::mbar_wait(&acc_full[buf], (cc >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc2::drain_add<16>(tmem_base + lane_off + buf * kNA + (uint32_t)(cs * 16), acc);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tcp::mbar_arrive_cluster_relaxed(acc_empty_leader + 8u * buf);
      }
      // ---- conv a result -> bias, ELU, split -> K-major operand of conv b: channels 16cs .. 16cs+15 of row `row` ----------
      {
        const uint32_t hi_a = smem_u + OFF_R + (uint32_t)((cs >> 1) * (kRHi / 2) + row * 128);
        const uint32_t lo_a = smem_u + OFF_R + kRHi + (uint32_t)((cs >> 1) * (kRLo / 2) + row * 64);
        const int k7 = row & 7, k3 = (row >> 1) & 3;
        uint32_t lob[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 h4, l4;
          split_tf32(elu_fast(acc[4 * q] + ba[4 * q]), h4.x, l4.x);
          split_tf32(elu_fast(acc[4 * q + 1] + ba[4 * q + 1]), h4.y, l4.y);
          split_tf32(elu_fast(acc[4 * q + 2] + ba[4 * q + 2]), h4.z, l4.z);
          split_tf32(elu_fast(acc[4 * q + 3] + ba[4 * q + 3]), h4.w, l4.w);
          sts128(hi_a + (uint32_t)((((cs & 1) * 4 + q) ^ k7) << 4), h4);
          lob[2 * q] = pack_bf16x2(l4.x, l4.y);
          lob[2 * q + 1] = pack_bf16x2(l4.z, l4.w);
        }
        // 16 bf16 = 32 B = two 16-byte pieces (cs%2)*2 + {0,1} of the 64-byte row, SWIZZLE_64B: piece ^= (row / 2) % 4
#pragma unroll
        for (int j = 0; j < 2; ++j)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lo_a + (uint32_t)((((cs & 1) * 2 + j) ^ k3) << 4)),
                       "r"(lob[4 * j]), "r"(lob[4 * j + 1]), "r"(lob[4 * j + 2]), "r"(lob[4 * j + 3]) : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tcp::mbar_arrive_cluster(r_ready_leader);
      // ---- conv b result -> the usual tile finish (bias, skip, ELU, split, stores) ------------------------------------------
      tc::mbar_wait(b_full, tc_ & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float accb[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) accb[i] = 0.f;
      tc2::drain_add<32>(tmem_base + lane_off + 128u + (uint32_t)(cs * 32), accb);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      if (mine) tc2::finish_tile<32, 16>(ep, accb, b, m0 + quarter * 32, cs * 32, Lout, stg, lane);
      ++tc_;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  tcp::cluster_sync();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

}  // namespace tcr
}  // namespace mimi
