// tcgen05 implicit-GEMM, fourth generation: activations live in HBM ONCE, as plain fp32 -- the ELU of the consuming
// conv and the TF32 hi/lo split happen inside the GEMM kernel, in shared memory, by a transform warpgroup that sits
// between the TMA producer and the MMA issuer. Against tc_gemm2 (producers store raw + hi + lo = 12 B per element,
// consumers load hi + lo = 8 B) every activation costs 4 B written and 4 B per tap read: the level-0/1 kernels that
// ran at 70-80 % of the HBM copy bandwidth move less than half the bytes, the workspace halves, and the epilogue
// shrinks to one store per element. The transform costs issue slots that were idle (the GEMM roles mostly wait on
// one another) and ~25 % of the shared-memory bank bandwidth.
//
//   16 warps: warpgroup 0 = {warp 0 TMA producer, warp 1 MMA issuer, 2 idle}, warpgroup 1 = transform (4 warps),
//   warpgroups 2-3 = epilogue (8 warps, the shared epilogue_role). setmaxnreg moves registers from the first two
//   warpgroups (72) to the epilogue (184).
//   stage = [A raw -> A_hi in place | A_lo | W_hi | W_lo]: TMA fills A (16 KB, raw fp32) and the pre-split weights
//   (static), `raw_full` fires, the transform warps rewrite A as hi (RN to 10 mantissa bits) and write lo next to it
//   (same swizzled positions, the transform is element-wise), `op_full` fires, the MMAs run (two per k-step, stacked
//   N as in tc_gemm2), `empty` returns the stage.
#pragma once
#include "tc_gemm2.cuh"

namespace mimi {
namespace tc4 {

using tc::Epilogue;
using tc::kBK;
using tc::kBM;
using tc::kChunkKB;
using tc::kUmmaK;
using tc2::Sched;

constexpr int kEpiWarps = 8;
constexpr int kEW0 = 8;                             // first epilogue warp
constexpr int kThreads = 512;
constexpr int kSmemMax = 232448;

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = kBM * kBK * 4;                      // 16 KB
  static constexpr int W_BYTES = BN * kBK * 4;
  static constexpr int STAGE = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int LOAD_BYTES = A_BYTES + 2 * W_BYTES;           // what TMA brings per stage
  static constexpr int PC = (BN / 2 >= 32) ? 32 : BN / 2;
  static constexpr int STG = kEpiWarps * 32 * PC * 4;
  static constexpr int BAR_BYTES = 512;
  static constexpr int STAGES_RAW = (kSmemMax - 1024 - STG - BAR_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = 1024 + STAGES * STAGE + STG + BAR_BYTES;
  static constexpr int TMEM_COLS = (4 * BN <= 32) ? 32 : (4 * BN <= 64) ? 64 : (4 * BN <= 128) ? 128 : (4 * BN <= 256) ? 256 : 512;
  static_assert(BN == 32 || BN == 64 || BN == 128, "BN");
  static_assert(STAGES >= 3, "ring too shallow");
};

// tmA: 3-D map over the raw fp32 activation (overlapping rows, as in tc_gemm2); elu_in: apply ELU while splitting
template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
tc4_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW_hi,
                const __grid_constant__ CUtensorMap tmW_lo, int K, int elu_in, const Epilogue ep, const Sched sc) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int STAGE = C::STAGE;
  constexpr int A_BYTES = C::A_BYTES;
  constexpr int W_BYTES = C::W_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + STAGES * STAGE;
  uint64_t* raw_full = reinterpret_cast<uint64_t*>(stg_base + C::STG);   // TMA landed (A raw + W hi/lo)
  uint64_t* op_full = raw_full + STAGES;                                  // A split in place (4 transform warps)
  uint64_t* empty_bar = op_full + STAGES;                                 // MMAs have read the stage
  uint64_t* acc_full = empty_bar + STAGES;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int nkb = K / kBK;
  const int ckb = ep.chunk_kb > 0 ? ep.chunk_kb : kChunkKB;
  const int nchunks = (nkb + ckb - 1) / ckb;
  const int vtiles = tc2::sched_tiles(sc) * sc.ntn;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA); tc::prefetch_tmap(&tmW_hi); tc::prefetch_tmap(&tmW_lo);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&raw_full[s], 1); tc::mbar_init(&op_full[s], 4); tc::mbar_init(&empty_bar[s], 1); }
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
    n0 = (id % sc.ntn) * BN;
    return tc2::sched_tile(sc, ep, id / sc.ntn, b, m0, Lout);
  };

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    if (warp == 0) {
      if (tc::elect_one()) {
        uint32_t kbc = 0;
        for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
          int b, m0, n0, Lout;
          if (!decode(id, b, m0, n0, Lout)) continue;
          for (int kb = 0; kb < nkb; ++kb, ++kbc) {
            const uint32_t s = kbc % STAGES;
            tc::mbar_wait(&empty_bar[s], ((kbc / STAGES) & 1u) ^ 1u);
            uint8_t* st = smem + s * STAGE;
            tc::mbar_expect_tx(&raw_full[s], C::LOAD_BYTES);
            const int kx = tc2::kblock_order(sc, kb) * kBK;
            tc::tma_load_3d(st, &tmA, &raw_full[s], kx, m0, b);
            tc::tma_load_2d(st + 2 * A_BYTES, &tmW_hi, &raw_full[s], kx, n0);
            tc::tma_load_2d(st + 2 * A_BYTES + W_BYTES, &tmW_lo, &raw_full[s], kx, n0);
          }
        }
      }
    } else if (warp == 1) {
      constexpr uint32_t idesc = tc::make_idesc(kBM, BN);
      constexpr uint32_t idesc2 = tc::make_idesc(kBM, 2 * BN);
      const uint32_t smem_base_u32 = tc::smem_u32(smem);
      uint32_t kbc = 0, cc = 0;
      for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
        int b, m0, n0, Lout;
        if (!decode(id, b, m0, n0, Lout)) continue;
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const uint32_t buf = cc & 1u;
          tc::mbar_wait(&acc_empty[buf], ((cc >> 1) & 1u) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_main = tmem_base + buf * (2 * BN);
          const int kb_end = min(nkb, (c + 1) * ckb);
          for (int kb = c * ckb; kb < kb_end; ++kb, ++kbc) {
            const uint32_t s = kbc % STAGES;
            tc::mbar_wait(&op_full[s], (kbc / STAGES) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d_ahi = tc::desc_lo(smem_base_u32 + s * STAGE);
            constexpr uint32_t kAlo = A_BYTES >> 4, kWhi = (2 * A_BYTES) >> 4;
            const bool first_in_chunk = kb == c * ckb;
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k) {
                tc::umma_tf32_lo(tmem_main, d_ahi + 2 * k, d_ahi + kWhi + 2 * k, idesc2, !(first_in_chunk && k == 0));
                tc::umma_tf32_lo(tmem_main + BN, d_ahi + kAlo + 2 * k, d_ahi + kWhi + 2 * k, idesc, 1u);
              }
              tc::umma_commit(&empty_bar[s]);
              if (kb + 1 == kb_end) tc::umma_commit(&acc_full[buf]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp < 8) {
    // ---- transform warpgroup: A raw -> (ELU) -> hi in place, lo next to it ------------------------------------------
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
    const int tt = threadIdx.x - 128;                // 0..127; chunk (16 B) index tt + 128*i, i = 0..7
    const uint32_t smem_base_u32 = tc::smem_u32(smem);
    uint32_t kbc = 0;
    for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
      int b, m0, n0, Lout;
      if (!decode(id, b, m0, n0, Lout)) continue;
      for (int kb = 0; kb < nkb; ++kb, ++kbc) {
        const uint32_t s = kbc % STAGES;
        tc::mbar_wait(&raw_full[s], (kbc / STAGES) & 1u);
        const uint32_t a0 = smem_base_u32 + s * STAGE + (uint32_t)tt * 16u;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = lds128(a0 + (uint32_t)i * 2048u);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 x = v[i];
          if (elu_in) { x.x = elu_fast(x.x); x.y = elu_fast(x.y); x.z = elu_fast(x.z); x.w = elu_fast(x.w); }
          float4 h4, l4;
          split_tf32(x.x, h4.x, l4.x); split_tf32(x.y, h4.y, l4.y); split_tf32(x.z, h4.z, l4.z); split_tf32(x.w, h4.w, l4.w);
          sts128(a0 + (uint32_t)i * 2048u, h4);
          sts128(a0 + (uint32_t)i * 2048u + A_BYTES, l4);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&op_full[s]);
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 184;");
    tc2::epilogue_role<BN, C::PC, kEpiWarps, kEW0>(ep, sc, stg_base, acc_full, acc_empty, tmem_base, nchunks, warp, lane);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
  }
}

// z [B][rows][512] -> replicate-padded copy for the stride-2 downsample conv (pad_mode="replicate",
// modeling_mimi.py:1422-1431): zp row 0,1 = z row 0; zp row 2+t = z row t; zp row 2+T = z row T-1. One warp per row.
__global__ void __launch_bounds__(256) pad_replicate_kernel(const float* __restrict__ z, long long z_item_stride,
                                                            float* __restrict__ zp, long long zp_item_stride,
                                                            const int* __restrict__ len, int uniform_len) {
  const int b = blockIdx.y;
  const int T = len ? len[b] : uniform_len;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;
  if (T <= 0 || r >= T + 3) return;
  const int src = min(max(r - 2, 0), T - 1);
  const float* zr = z + (long long)b * z_item_stride + (long long)src * kHidden;
  float* o = zp + (long long)b * zp_item_stride + (long long)r * kHidden;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    *reinterpret_cast<float4*>(o + c) = ld_nc_f4(zr + c);
  }
}

// zero the halo rows of a raw buffer: rows [0, front) and [front + L_b, front + L_b + back) of every item
__global__ void zero_halo_raw_kernel(float* x, long long item_stride, int C, int front, int back,
                                     const int* __restrict__ len, int uniform_len) {
  const int b = blockIdx.y;
  const int L = len ? len[b] : uniform_len;
  const int per = (front + back) * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
    int r = i / C;
    const int c = i - r * C;
    if (r >= front) r = front + L + (r - front);
    x[(long long)b * item_stride + (long long)r * C + c] = 0.f;
  }
}

}  // namespace tc4
}  // namespace mimi
