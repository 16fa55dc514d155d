// tcgen05 implicit-GEMM, third generation (EXPERIMENTAL, debug_set mode 4; not the default): cuts the L2 -> shared-memory operand stream, which is what bounds the
// 3xTF32 GEMMs on B200 (8 bytes per operand element; tc_gemm2's 128x128 tiles need 85 B/clk/SM against a
// measured ~35 B/clk/SM delivered, profiles/r01_tc2_gemm_ncu.md).
//
//   * 256 x BN output tile per CTA as two 128-row halves that share every weight k-block (24 MMAs per 32 KB of
//     weights instead of 12): weight bytes per output row halve.
//   * No im2col redundancy: a conv with kernel k = G*stride reads each input row ONCE. The activation is staged
//     as "planes": plane (ph, p) = input rows {q*stride + ph} x channels [32p, 32p+32), 128 B per row, through a
//     4-D TMA map; tap tau = ph + stride*dq is the SAME plane with the UMMA descriptor start shifted by dq rows
//     (legal with SWIZZLE_128B because the swizzle is a function of absolute smem address bits, see
//     profiles/r01_sw128_row_shift_probe.md). Activation bytes drop by G (2 for the strided convs, 3 for k=3).
//   * One accumulator per half and chunk: the cross terms hi*lo + lo*hi go into the main accumulator, which is
//     drained into fp32 registers every `chunk_kb` k-blocks (default 2, K = 64: 24 tensor-core adds per drain;
//     measured 6e-7 relative error per GEMM, tools/gpu_acc_experiment.py). TMEM = 2 halves x 2 buffers x BN.
//
// Status (profiles/r01_tc3_notes.md): correct (unit + pipeline parity), same speed as tc_gemm2 -- the 2 + 2 stage
// rings that fit next to two 66 KB activation planes are too shallow to cover the L2 latency, so the halved
// operand stream does not turn into time. Kept for the next round (deeper rings need smaller planes).
//
// Roles: warp 0 TMA producer (activation-plane ring + weight ring), warp 1 MMA issuer, warps 2-3 idle (they
// complete warpgroup 0, which hands its registers over with setmaxnreg: the 128 + 32 live fp32 values per epilogue
// thread do not fit the 168 registers a 384-thread CTA starts with), warps 4..11 epilogue (chunk drains, then
// bias / GELU / LayerScale, staged transposes, coalesced stores).
#pragma once
#include "tc_gemm2.cuh"

namespace mimi {
namespace tc3 {

using tc::Epilogue;
using tc::kBK;
using tc::kUmmaK;

constexpr int kBM = 256;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 128 + 32 * kEpiWarps;      // 384: warpgroup 0 = {TMA, MMA, 2 idle warps}, warpgroups 1-2 = epilogue
constexpr int kARows = 264;                         // 256 + G - 1 <= 258 rows per plane, rounded to whole KB
constexpr int kAHalf = kARows * 128;                // bytes per hi (or lo) plane
constexpr int kAStage = 2 * kAHalf;
constexpr int kAStages = 2;
constexpr int kSmemMax = 232448;

struct Geom {
  int G;          // taps per plane = k / stride (1, 2 or 3)
  int s;          // conv stride
  int cpanels;    // C_in / 32
  int B;          // items
  int mt_max;     // 256-row tiles per item at the longest item
  int ntn;        // N / BN
  int chunk_kb;   // k-blocks per accumulation chunk
};

template <int BN>
struct Cfg {
  static constexpr int W_BYTES = BN * kBK * 4;
  static constexpr int W_STAGE = 2 * W_BYTES;
  static constexpr int PC = 16;                                      // staging piece width (columns)
  static constexpr int STG_WARP = 32 * PC * 4;
  static constexpr int STG = kEpiWarps * STG_WARP;                   // 16 KB
  static constexpr int BAR_BYTES = 512;
  static constexpr int W_STAGES_RAW = (kSmemMax - 1024 - kAStages * kAStage - STG - BAR_BYTES) / W_STAGE;
  static constexpr int W_STAGES = W_STAGES_RAW > 6 ? 6 : W_STAGES_RAW;
  static constexpr int SMEM = 1024 + kAStages * kAStage + W_STAGES * W_STAGE + STG + BAR_BYTES;
  static constexpr int TMEM_COLS = (4 * BN <= 256) ? 256 : 512;
  static_assert(BN == 64 || BN == 128, "BN");
  static_assert(W_STAGES >= 2, "weight ring too shallow");
};

using tc2::tma_load_4d;

// tmA_*: 4-D maps {C, stride, q, item}, box {32, 1, 128 (+G-1 for the second box), 1}: see tc_host.inl
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc3_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmA2_hi, const __grid_constant__ CUtensorMap tmA2_lo,
                const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                const Epilogue ep, const Geom gm) {
  using C = Cfg<BN>;
  constexpr int W_STAGES = C::W_STAGES;
  constexpr int W_STAGE = C::W_STAGE;
  constexpr int W_BYTES = C::W_BYTES;
  constexpr int HALF = BN / 2;                     // columns per epilogue thread
  constexpr int PC = C::PC;
  constexpr int NP = HALF / PC;
  constexpr int LPR = PC / 4;                      // 4 lanes per row
  constexpr int RPI = 32 / LPR;                    // 8 rows per warp instruction
  constexpr int IT = 32 / RPI;                     // 4 instructions per 32-row piece

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_base = smem + kAStages * kAStage;
  uint8_t* stg_base = w_base + W_STAGES * W_STAGE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(stg_base + C::STG);
  uint64_t* a_empty = a_full + kAStages;
  uint64_t* w_full = a_empty + kAStages;
  uint64_t* w_empty = w_full + W_STAGES;
  uint64_t* acc_full = w_empty + W_STAGES;         // [2]
  uint64_t* acc_empty = acc_full + 2;              // [2]
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nplanes = gm.s * gm.cpanels;
  const int nkb = nplanes * gm.G;
  const int ckb = gm.chunk_kb > 0 ? gm.chunk_kb : 2;
  const int vtiles = gm.mt_max * gm.B * gm.ntn;
  const int rows2 = 128 + gm.G - 1;                // rows of the second activation box
  const uint32_t a_bytes = (uint32_t)(2 * (128 + rows2) * 128);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi); tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmA2_hi); tc::prefetch_tmap(&tmA2_lo);
    tc::prefetch_tmap(&tmW_hi); tc::prefetch_tmap(&tmW_lo);
    for (int s = 0; s < kAStages; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < W_STAGES; ++s) { tc::mbar_init(&w_full[s], 1); tc::mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], kEpiWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(C::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;

  auto decode = [&](int id, int& b, int& m0, int& n0, int& Lout) {
    const int nt = id % gm.ntn;
    const int t = id / gm.ntn;
    b = t % gm.B;
    m0 = (t / gm.B) * kBM;
    n0 = nt * BN;
    const int Lin = ep.len_in ? __ldg(ep.len_in + b) : ep.uniform_len_in;
    Lout = (Lin + ep.conv_stride - 1) / ep.conv_stride;
    return m0 < Lout;
  };

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
  if (warp == 0) {
    if (lane == 0) {
      uint32_t ac = 0, wc = 0;                     // plane / weight-block counters (ring positions)
      for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
        int b, m0, n0, Lout;
        if (!decode(id, b, m0, n0, Lout)) continue;
        for (int pi = 0; pi < nplanes; ++pi, ++ac) {
          const int ph = pi / gm.cpanels, pn = pi - ph * gm.cpanels;
          const uint32_t as = ac % kAStages;
          tc::mbar_wait(&a_empty[as], ((ac / kAStages) & 1u) ^ 1u);
          uint8_t* st = smem + as * kAStage;
          tc::mbar_expect_tx(&a_full[as], a_bytes);
          tma_load_4d(st, &tmA_hi, &a_full[as], pn * 32, ph, m0, b);
          tma_load_4d(st + 128 * 128, &tmA2_hi, &a_full[as], pn * 32, ph, m0 + 128, b);
          tma_load_4d(st + kAHalf, &tmA_lo, &a_full[as], pn * 32, ph, m0, b);
          tma_load_4d(st + kAHalf + 128 * 128, &tmA2_lo, &a_full[as], pn * 32, ph, m0 + 128, b);
          for (int dq = 0; dq < gm.G; ++dq, ++wc) {
            const int kbw = (ph + gm.s * dq) * gm.cpanels + pn;     // weight k-block of tap ph + s*dq, panel pn
            const uint32_t ws = wc % W_STAGES;
            tc::mbar_wait(&w_empty[ws], ((wc / W_STAGES) & 1u) ^ 1u);
            uint8_t* wt = w_base + ws * W_STAGE;
            tc::mbar_expect_tx(&w_full[ws], W_STAGE);
            tc::tma_load_2d(wt, &tmW_hi, &w_full[ws], kbw * kBK, n0);
            tc::tma_load_2d(wt + W_BYTES, &tmW_lo, &w_full[ws], kbw * kBK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::make_idesc(128, BN);
      uint32_t ac = 0, wc = 0, cc = 0;
      for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
        int b, m0, n0, Lout;
        if (!decode(id, b, m0, n0, Lout)) continue;
        const int nhalves = (m0 + 128 < Lout) ? 2 : 1;
        int kb = 0;                                 // k-blocks consumed in this tile
        for (int pi = 0; pi < nplanes; ++pi, ++ac) {
          const uint32_t as = ac % kAStages;
          tc::mbar_wait(&a_full[as], (ac / kAStages) & 1u);
          const uint32_t a_hi0 = tc::smem_u32(smem + as * kAStage);
          for (int dq = 0; dq < gm.G; ++dq, ++wc, ++kb) {
            const uint32_t buf = cc & 1u;
            if (kb % ckb == 0) tc::mbar_wait(&acc_empty[buf], ((cc >> 1) & 1u) ^ 1u);   // drained two chunks ago
            const uint32_t ws = wc % W_STAGES;
            tc::mbar_wait(&w_full[ws], (wc / W_STAGES) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t w_hi = tc::smem_u32(w_base + ws * W_STAGE);
            const uint32_t w_lo = w_hi + W_BYTES;
            const bool first_in_chunk = (kb % ckb) == 0;
            for (int hf = 0; hf < nhalves; ++hf) {
              const uint32_t a_hi = a_hi0 + (uint32_t)((hf * 128 + dq) * 128);
              const uint32_t a_lo = a_hi + kAHalf;
              const uint32_t tmem_acc = tmem_base + (uint32_t)(hf * 2 + buf) * BN;
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k)
                tc::umma_tf32(tmem_acc, tc::make_smem_desc(a_hi + k * 32), tc::make_smem_desc(w_hi + k * 32), idesc,
                              !(first_in_chunk && k == 0));
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k)
                tc::umma_tf32(tmem_acc, tc::make_smem_desc(a_lo + k * 32), tc::make_smem_desc(w_hi + k * 32), idesc, 1u);
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k)
                tc::umma_tf32(tmem_acc, tc::make_smem_desc(a_hi + k * 32), tc::make_smem_desc(w_lo + k * 32), idesc, 1u);
            }
            tc::umma_commit(&w_empty[ws]);
            if ((kb + 1) % ckb == 0 || kb + 1 == nkb) { tc::umma_commit(&acc_full[buf]); ++cc; }
          }
          tc::umma_commit(&a_empty[as]);
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ---- epilogue warps ------------------------------------------------------------------------------------
    const int ew = warp - 4;
    const int quarter = warp & 3;
    const int half = ew >> 2;
    const int col0 = half * HALF;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t stg = tc::smem_u32(stg_base + ew * C::STG_WARP);     // [32 rows][4 float4], swizzled
    const int rr = lane / LPR;
    const int cj = lane % LPR;
    const int nchunks = (nkb + ckb - 1) / ckb;
    uint32_t cc = 0;
    for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
      int b, m0, n0, Lout;
      if (!decode(id, b, m0, n0, Lout)) continue;
      const int nhalves = (m0 + 128 < Lout) ? 2 : 1;
      float acc[2][HALF];
#pragma unroll
      for (int hf = 0; hf < 2; ++hf)
#pragma unroll
        for (int i = 0; i < HALF; ++i) acc[hf][i] = 0.f;
      for (int c = 0; c < nchunks; ++c, ++cc) {
        const uint32_t buf = cc & 1u;
        tc::mbar_wait(&acc_full[buf], (cc >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc2::drain_add<HALF>(tmem_base + lane_off + buf * BN + (uint32_t)col0, acc[0]);
        if (nhalves == 2) tc2::drain_add<HALF>(tmem_base + lane_off + (2u + buf) * BN + (uint32_t)col0, acc[1]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
      }
      const int ncol0 = n0 + col0;
      const long long raw_base = (long long)b * ep.raw_item_stride + ncol0;
      const long long split_base = (long long)b * ep.split_item_stride + (long long)ep.split_front * ep.N + ncol0;
      // One copy of the epilogue math, looped at run time over (half, 16-column piece): fully unrolled it is
      // ~12k instructions (190 KB), far beyond the instruction caches, and instruction fetch became the bottleneck
      // (profiles/r01_tc3_icache.md). The piece is first selected into a small register array.
#pragma unroll 1
      for (int hp = 0; hp < nhalves * NP; ++hp) {
        const int hf = hp / NP, p = hp - hf * NP;
        float tmp[PC];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2)
#pragma unroll
          for (int p2 = 0; p2 < NP; ++p2)
            if (h2 == hf && p2 == p) {
#pragma unroll
              for (int i = 0; i < PC; ++i) tmp[i] = acc[h2][p2 * PC + i];
            }
        const int row_base = m0 + hf * 128 + quarter * 32;
        float4 resv[IT];
        if (ep.res) {
#pragma unroll
          for (int it = 0; it < IT; ++it) {
            const int row = row_base + it * RPI + rr;
            resv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < Lout) resv[it] = *reinterpret_cast<const float4*>(ep.res + raw_base + (long long)row * ep.N + p * PC + cj * 4);
          }
        }
        const int wkey = (lane >> 1) & (LPR - 1);
#pragma unroll
        for (int j = 0; j < LPR; ++j) {
          float4 v = make_float4(tmp[4 * j], tmp[4 * j + 1], tmp[4 * j + 2], tmp[4 * j + 3]);
          const int c = ncol0 + p * PC + 4 * j;
          if (ep.bias) {
            const float4 t = ld_nc_f4(ep.bias + c);
            v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
          }
          if (ep.act == 1) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
          if (ep.scale) {
            const float4 t = ld_nc_f4(ep.scale + c);
            v.x *= t.x; v.y *= t.y; v.z *= t.z; v.w *= t.w;
          }
          sts128(stg + (uint32_t)(lane * LPR + (j ^ wkey)) * 16u, v);
        }
        __syncwarp();
        float4 tv[IT];
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const int r = it * RPI + rr;
          const int rkey = (r >> 1) & (LPR - 1);
          tv[it] = lds128(stg + (uint32_t)(r * LPR + (cj ^ rkey)) * 16u);
        }
#pragma unroll
        for (int it = 0; it < IT; ++it) {
          const int r = it * RPI + rr;
          float4 v = tv[it];
          const int row = row_base + r;
          if (row < Lout) {
            const long long o = (long long)row * ep.N + p * PC + cj * 4;
            if (ep.res) { v.x += resv[it].x; v.y += resv[it].y; v.z += resv[it].z; v.w += resv[it].w; }
            if (ep.out_raw) *reinterpret_cast<float4*>(ep.out_raw + raw_base + o) = v;
            if (ep.out_hi) {
              if (ep.elu_split) { v.x = elu_fast(v.x); v.y = elu_fast(v.y); v.z = elu_fast(v.z); v.w = elu_fast(v.w); }
              store_split4(ep.out_hi + split_base + o, ep.out_lo + split_base + o, v);
            }
          }
        }
        __syncwarp();
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
  }
}

}  // namespace tc3
}  // namespace mimi
