// Split residual vector quantiser on tensor cores (MimiSplitResidualVectorQuantizer.encode, modeling_mimi.py:1311-1338
// over MimiResidualVectorQuantizer.encode :1262-1280 and MimiEuclideanCodebook.quantize :1197-1202).
//
// One CTA owns 64 frames and walks all K stages. Per stage the score matrix is computed TRANSPOSED,
//     D[code, frame] = E[code, :] . r[frame, :]        (M = 128 codes per block, N = 64 frames, K = 256)
// so that the streamed operand (the 2048 x 256 codebook, 16 blocks of 128 codes through a TMA ring) is the
// M = 128 side and the tensor core runs at full rate, while the 64 x 256 residual stays resident in shared memory as
// the B operand (TF32 hi | lo stacked per k-block, SWIZZLE_128B, written by the threads themselves). 3xTF32:
// one N = 128 MMA gives E_hi r_hi^T (columns 0..63) and E_hi r_lo^T (columns 64..127), a second N = 64 MMA adds
// E_lo r_hi^T to columns 64..127. The epilogue threads (TMEM lane = code) turn each block into the torch.cdist
// distance sqrt(max(|r|^2 + |e|^2 - 2 e.r, 0)) and keep a running (min, lowest index) per frame in registers: the
// 2048-wide distance rows never exist anywhere. After 16 blocks a shuffle + shared-memory reduction picks the code of
// every frame, the residual update r -= E[idx] is applied in fp32 (hi + lo is exact) and the operand is re-split.
#pragma once
#include "tc_gemm.cuh"

namespace mimi {
namespace rvqtc {

constexpr int kFrames = 64;                        // frames per CTA (MMA N)
constexpr int kCodesPerBlock = 128;                // MMA M
constexpr int kBlocks = kCodebookSize / kCodesPerBlock;     // 16
constexpr int kKB = kCodeDim / 32;                 // 8 k-blocks
constexpr int kRKb = 2 * kFrames * 128;            // bytes of one residual k-block: hi (64 rows) | lo (64 rows) = 16 KB
constexpr int kRBytes = kKB * kRKb;                // 128 KB
constexpr int kAStage = 2 * kCodesPerBlock * 128;  // codebook stage: hi | lo = 32 KB
constexpr int kAStages = 2;
constexpr int kThreads = 320;
constexpr int kMisc = 4096;                        // xn[64], candidates, barriers
constexpr int kSmem = 1024 + kRBytes + kAStages * kAStage + kMisc;

struct Params {
  const float* rproj;        // [B][item_stride]: row t = [P_sem e (256) | P_aco e (256)]
  long long item_stride;
  const float* embed;        // [32][2048][256] fp32 row-major (gather for the residual update)
  const float* enorm;        // [32][2048] |e|^2
  long long* codes;          // [B][K][T_out] int64
  int K, T_out;
  const int* len;            // device [B] frames per item or nullptr -> uniform_len
  int uniform_len;
  int B;
  int total_frames;
  const int* frame_prefix;   // device [B+1] prefix sums of len (ragged) or nullptr
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// tmE_hi / tmE_lo: 2-D maps over the split codebooks [32 * 2048 rows][256], box {32, 128}, SWIZZLE_128B
__global__ void __launch_bounds__(kThreads, 1)
rvq_tc_kernel(const __grid_constant__ CUtensorMap tmE_hi, const __grid_constant__ CUtensorMap tmE_lo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* r_op = smem;                                      // residual operand: 8 x (hi 8 KB | lo 8 KB)
  uint8_t* a_ring = smem + kRBytes;
  uint8_t* misc = a_ring + kAStages * kAStage;
  float* xn = reinterpret_cast<float*>(misc);                // [64] |r|^2
  float* cand_d = xn + kFrames;                              // [4 quarters][64 frames]
  int* cand_i = reinterpret_cast<int*>(cand_d + 4 * kFrames);
  int* best = cand_i + 4 * kFrames;                          // [64] chosen code
  int* fr_b = best + kFrames;                                // [64] item of frame
  int* fr_t = fr_b + kFrames;                                // [64] frame index inside the item
  uint64_t* bars = reinterpret_cast<uint64_t*>(fr_t + kFrames);
  uint64_t* full_bar = bars;                                 // [kAStages]
  uint64_t* empty_bar = bars + kAStages;                     // [kAStages]
  uint64_t* acc_full = bars + 2 * kAStages;                  // [2]
  uint64_t* acc_empty = acc_full + 2;                        // [2]
  uint64_t* r_ready = acc_empty + 2;                         // residual operand of this stage staged (8 warps)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(r_ready + 1);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int f0 = blockIdx.x * kFrames;
  const int nf = min(kFrames, p.total_frames - f0);

  if (threadIdx.x == 0) {
    tc::prefetch_tmap(&tmE_hi); tc::prefetch_tmap(&tmE_lo);
    for (int s = 0; s < kAStages; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { tc::mbar_init(&acc_full[s], 1); tc::mbar_init(&acc_empty[s], 8); }
    tc::mbar_init(r_ready, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (threadIdx.x >= 64 && threadIdx.x < 64 + kFrames) {
    const int f = threadIdx.x - 64;
    int b = 0, t = 0;
    const int m = f0 + f;
    if (f < nf) {
      if (p.frame_prefix) {
        int lo = 0, hi = p.B;                        // largest b with prefix[b] <= m
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (__ldg(p.frame_prefix + mid) <= m) lo = mid; else hi = mid;
        }
        b = lo; t = m - __ldg(p.frame_prefix + lo);
      } else {
        b = m / p.uniform_len; t = m - b * p.uniform_len;
      }
    }
    fr_b[f] = b; fr_t[f] = t;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t c = 0;
      for (int stage = 0; stage < p.K; ++stage)
        for (int blk = 0; blk < kBlocks; ++blk)
          for (int kb = 0; kb < kKB; ++kb, ++c) {
            const uint32_t s = c % kAStages;
            tc::mbar_wait(&empty_bar[s], ((c / kAStages) & 1u) ^ 1u);
            uint8_t* st = a_ring + s * kAStage;
            tc::mbar_expect_tx(&full_bar[s], kAStage);
            tc::tma_load_2d(st, &tmE_hi, &full_bar[s], kb * 32, stage * kCodebookSize + blk * kCodesPerBlock);
            tc::tma_load_2d(st + kAStage / 2, &tmE_lo, &full_bar[s], kb * 32, stage * kCodebookSize + blk * kCodesPerBlock);
          }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc128 = tc::make_idesc(128, 128);
      constexpr uint32_t idesc64 = tc::make_idesc(128, 64);
      const uint32_t d_rop = tc::desc_lo(tc::smem_u32(r_op));
      uint32_t c = 0, bc = 0;
      for (int stage = 0; stage < p.K; ++stage) {
        tc::mbar_wait(r_ready, (uint32_t)stage & 1u);
        for (int blk = 0; blk < kBlocks; ++blk, ++bc) {
          const uint32_t buf = bc & 1u;
          tc::mbar_wait(&acc_empty[buf], ((bc >> 1) & 1u) ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t acc = tmem_base + buf * 128;          // columns [0,64) main | [64,128) cross terms
          for (int kb = 0; kb < kKB; ++kb, ++c) {
            const uint32_t s = c % kAStages;
            tc::mbar_wait(&full_bar[s], (c / kAStages) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t e_hi = tc::desc_lo(tc::smem_u32(a_ring) + s * kAStage);
            const uint32_t e_lo = e_hi + ((kAStage / 2) >> 4);
            const uint32_t r_st = d_rop + ((kb * kRKb) >> 4);   // [r_hi (64 rows) | r_lo (64 rows)] stacked: N = 128
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                tc::umma_tf32_lo(acc, e_hi + 2 * k, r_st + 2 * k, idesc128, (uint32_t)((kb | k) != 0));
                tc::umma_tf32_lo(acc + 64, e_lo + 2 * k, r_st + 2 * k, idesc64, 1u);
              }
              tc::umma_commit(&empty_bar[s]);
              if (kb + 1 == kKB) tc::umma_commit(&acc_full[buf]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ---- 8 epilogue warps: TMEM lane quarter = warp & 3 (codes), frame half = (warp - 2) >> 2 ----------------------
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int fh = ew >> 2;                        // frames [32*fh, 32*fh + 32)
    const int et = threadIdx.x - 64;               // 0..255
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t rop = tc::smem_u32(r_op);
    auto epi_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };
    // residual ownership for loads / updates: thread et handles frame uf = et / 4, dims [64*(et%4), +64)
    const int uf = et >> 2, uq = et & 3;
    uint32_t bc = 0;
    for (int stage = 0; stage < p.K; ++stage) {
      // ---- stage the residual operand: stages 0 / 1 start from the semantic / acoustic projection --------------
      float part = 0.f;
      if (stage <= 1) {
        const float* src = p.rproj + (long long)fr_b[uf] * p.item_stride + (long long)fr_t[uf] * 512 + stage * 256 + uq * 64;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (uf < nf) v = ld_nc_f4(src + j * 4);
          part = fmaf(v.x, v.x, part); part = fmaf(v.y, v.y, part); part = fmaf(v.z, v.z, part); part = fmaf(v.w, v.w, part);
          float4 h4, l4;
          split_tf32(v.x, h4.x, l4.x); split_tf32(v.y, h4.y, l4.y); split_tf32(v.z, h4.z, l4.z); split_tf32(v.w, h4.w, l4.w);
          const int d = uq * 64 + j * 4;
          const uint32_t a = rop + (uint32_t)((d >> 5) * kRKb + uf * 128 + ((((d & 31) >> 2) ^ (uf & 7)) << 4));
          sts128(a, h4);
          sts128(a + kFrames * 128, l4);
        }
      } else {
        // r -= E[best] (modeling_mimi.py:1277), exact fp32 on hi + lo, then re-split
        const float* e = p.embed + ((long long)(stage - 1) * kCodebookSize + best[uf]) * kCodeDim + uq * 64;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int d = uq * 64 + j * 4;
          const uint32_t a = rop + (uint32_t)((d >> 5) * kRKb + uf * 128 + ((((d & 31) >> 2) ^ (uf & 7)) << 4));
          const float4 h0 = lds128(a), l0 = lds128(a + kFrames * 128);
          const float4 ev = ld_nc_f4(e + j * 4);
          float4 v = make_float4((h0.x + l0.x) - ev.x, (h0.y + l0.y) - ev.y, (h0.z + l0.z) - ev.z, (h0.w + l0.w) - ev.w);
          part = fmaf(v.x, v.x, part); part = fmaf(v.y, v.y, part); part = fmaf(v.z, v.z, part); part = fmaf(v.w, v.w, part);
          float4 h4, l4;
          split_tf32(v.x, h4.x, l4.x); split_tf32(v.y, h4.y, l4.y); split_tf32(v.z, h4.z, l4.z); split_tf32(v.w, h4.w, l4.w);
          sts128(a, h4);
          sts128(a + kFrames * 128, l4);
        }
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      if (uq == 0) xn[uf] = part;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      epi_sync();                                   // operand + xn complete for all frames
      if (lane == 0) tc::mbar_arrive(r_ready);

      // ---- 16 code blocks: distances and running (min, lowest index) for this thread's code lane ---------------
      float bd[32];
      int bi[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { bd[i] = INFINITY; bi[i] = 0; }
      const float* en = p.enorm + (long long)stage * kCodebookSize + quarter * 32 + lane;
      for (int blk = 0; blk < kBlocks; ++blk, ++bc) {
        const uint32_t buf = bc & 1u;
        const float e2 = __ldg(en + blk * kCodesPerBlock);
        const int code = blk * kCodesPerBlock + quarter * 32 + lane;
        tc::mbar_wait(&acc_full[buf], (bc >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t rm[32], rs[32];
        tmem_ld32(tmem_base + lane_off + buf * 128 + fh * 32, rm);
        tmem_ld32(tmem_base + lane_off + buf * 128 + 64 + fh * 32, rs);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float dot = __uint_as_float(rm[i]) + __uint_as_float(rs[i]);
          const float d2 = fmaf(-2.f, dot, xn[fh * 32 + i] + e2);
          const float d = sqrtf(fmaxf(d2, 0.f));
          if (d < bd[i]) { bd[i] = d; bi[i] = code; }      // codes ascend per thread: strict '<' keeps the lowest index
        }
      }
      // ---- argmin over the 128 code lanes x 16 blocks: warp shuffles, then the 4 quarters through smem -----------
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        float d = bd[i];
        int c = bi[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float d2 = __shfl_xor_sync(0xffffffffu, d, o);
          const int c2 = __shfl_xor_sync(0xffffffffu, c, o);
          if (d2 < d || (d2 == d && c2 < c)) { d = d2; c = c2; }
        }
        if (lane == 0) { cand_d[quarter * kFrames + fh * 32 + i] = d; cand_i[quarter * kFrames + fh * 32 + i] = c; }
      }
      epi_sync();
      if (et < kFrames) {
        float d = cand_d[et];
        int c = cand_i[et];
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          const float d2 = cand_d[q * kFrames + et];
          const int c2 = cand_i[q * kFrames + et];
          if (d2 < d || (d2 == d && c2 < c)) { d = d2; c = c2; }
        }
        best[et] = c;
        if (et < nf) p.codes[((long long)fr_b[et] * p.K + stage) * p.T_out + fr_t[et]] = (long long)c;
      }
      epi_sync();                                   // best[] visible; cand arrays free for the next stage
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

}  // namespace rvqtc
}  // namespace mimi
