// Split residual vector quantiser on tensor cores, fp16-pair generation (mode 9). Same structure and the same decisions as
// rvq_tc.cuh (MimiSplitResidualVectorQuantizer.encode, modeling_mimi.py:1311-1338; MimiEuclideanCodebook.quantize :1197-1202):
// one CTA = 64 frames x all K stages, transposed score GEMM D[code, frame] = E r^T with the codebook as the streamed M = 128
// side, torch.cdist's sqrt(max(|r|^2 + |e|^2 - 2 e.r, 0)), lowest index among equal minima, exact fp32 residual update.
//
// What changed: the kernel is bound by the codebook stream (every CTA pulls all K x 2048 x 256 entries through L2 -> shared
// memory; with TF32 hi / lo that is 4 MB per stage per CTA, 4 GB per launch at K = 8). Here both operands are fp16 pairs like
// every other GEMM operand of mode 9 (common.cuh: split_f16):
//   codebook   rows scaled by a power of two (max |e'| in [2^13, 2^14)), e' = e_hi + e_lo: 4 bytes per entry instead of 8;
//   residual   r = r_hi + r_lo / 2048, re-split from the fp32 residual every stage.
// Per K = 16 step: one N = 128 MMA  e_hi x [r_hi | r_lo]  (columns 0..63 main, 64..127 the term that carries the 1/2048) and
// one N = 64 MMA  e_lo x r_hi  into the main columns; dot = (main + cross / 2048) * 2^-s_code, the unscaling folded into the
// -2 of the distance. Half the bytes per stage also means a ring of FOUR 32 KB stages (a whole 128-code block in flight)
// where the TF32 kernel had room for two. The fp32 residual itself lives in global memory (in place in the acoustic half of
// the projection buffer, each thread re-reading only what it wrote): hi + lo / 2048 is 22 bits, not the exact value.
#pragma once
#include "front_f16.cuh"
#include "rvq_tc.cuh"

namespace mimi {
namespace rvq16 {

constexpr int kFrames = 64;                        // frames per CTA (MMA N)
constexpr int kCodesPerBlock = 128;                // MMA M
constexpr int kBlocks = kCodebookSize / kCodesPerBlock;     // 16
constexpr int kKB = kCodeDim / 64;                 // 4 k-blocks of 64 halfs (one 128-byte swizzle row)
constexpr int kRKb = 2 * kFrames * 128;            // residual k-block: hi (64 rows) | lo (64 rows) = 16 KB
constexpr int kRBytes = kKB * kRKb;                // 64 KB
constexpr int kAStage = 2 * kCodesPerBlock * 128;  // codebook stage: hi | lo = 32 KB
constexpr int kAStages = 4;
constexpr int kThreads = 320;
constexpr int kMisc = 4096;                        // xn[64], candidates, barriers
constexpr float kBandUp = 1.0f + 4.76837158203125e-07f;   // 1 + 2^-21, see the block loop
constexpr int kSmem = 1024 + kRBytes + kAStages * kAStage + kMisc;

struct Params {
  float* rproj;              // [B][item_stride]: row t = [P_sem e (256) | P_aco e (256)]; the acoustic half becomes the residual
  long long item_stride;
  const float* embed;        // [32][2048][256] fp32 row-major (gather for the residual update)
  const float* enorm;        // [32][2048] |e|^2
  const float* m2s;          // [32][2048] -2 * 2^-s: the distance's -2 times the unscaling of the code's fp16 row
  long long* codes;          // [B][K][T_out] int64
  int K, T_out;
  const int* len;            // device [B] frames per item or nullptr -> uniform_len
  int uniform_len;
  int B;
  int total_frames;
  const int* frame_prefix;   // device [B+1] prefix sums of len (ragged) or nullptr
};

// tmE_hi / tmE_lo: 2-D maps over the fp16 codebooks [32 * 2048 rows][256], box {64, 128}, SWIZZLE_128B
__global__ void __launch_bounds__(kThreads, 1)
rvq_f16_kernel(const __grid_constant__ CUtensorMap tmE_hi, const __grid_constant__ CUtensorMap tmE_lo, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* r_op = smem;                                      // residual operand: 4 x (hi 8 KB | lo 8 KB)
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
            tc::tma_load_2d(st, &tmE_hi, &full_bar[s], kb * 64, stage * kCodebookSize + blk * kCodesPerBlock);
            tc::tma_load_2d(st + kAStage / 2, &tmE_lo, &full_bar[s], kb * 64, stage * kCodebookSize + blk * kCodesPerBlock);
          }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc128 = tcp::make_idesc_f16(128, 128);
    constexpr uint32_t idesc64 = tcp::make_idesc_f16(128, 64);
    const uint32_t d_rop = tc::desc_lo(tc::smem_u32(r_op));
    uint32_t c = 0, bc = 0;
    for (int stage = 0; stage < p.K; ++stage) {
      tc::mbar_wait(r_ready, (uint32_t)stage & 1u);
      for (int blk = 0; blk < kBlocks; ++blk, ++bc) {
        const uint32_t buf = bc & 1u;
        tc::mbar_wait(&acc_empty[buf], ((bc >> 1) & 1u) ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = tmem_base + buf * 128;          // columns [0,64) main | [64,128) e_hi x r_lo (carries 1/2048)
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
              f1::umma_f16(acc, e_hi + 2 * k, r_st + 2 * k, idesc128, (uint32_t)((kb | k) != 0));
              f1::umma_f16(acc, e_lo + 2 * k, r_st + 2 * k, idesc64, 1u);
            }
            tc::umma_commit(&empty_bar[s]);
            if (kb + 1 == kKB) tc::umma_commit(&acc_full[buf]);
          }
          __syncwarp();
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
    // residual ownership for loads / updates: thread et handles frame uf = et / 4, dims [64 uq, +64) = k-block uq of its row
    const int uf = et >> 2, uq = et & 3;
    const uint32_t rrow = rop + (uint32_t)(uq * kRKb + uf * 128);
    float* rsd = p.rproj + (long long)fr_b[uf] * p.item_stride + (long long)fr_t[uf] * 512 + uq * 64;   // + 256: acoustic half
    __half2 mx2 = __floats2half2_rn(0.f, 0.f);     // running max |r_hi| (fp16 range check)
    uint32_t bc = 0;
    for (int stage = 0; stage < p.K; ++stage) {
      // ---- stage the residual operand: stages 0 / 1 start from the semantic / acoustic projection; later stages subtract
      //      the previous stage's code vector in fp32 (r -= E[best], modeling_mimi.py:1277) and keep the result in place ----
      float part = 0.f;
      const float* e = stage >= 2 ? p.embed + ((long long)(stage - 1) * kCodebookSize + best[uf]) * kCodeDim + uq * 64 : nullptr;
      float* rp = rsd + (stage >= 1 ? 256 : 0);
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = (jj + 2 * uq) & 7;            // the four threads of a frame write four different bank groups
        float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
        if (uf < nf) {
          if (stage <= 1) {
            v0 = ld_nc_f4(rp + j * 8); v1 = ld_nc_f4(rp + j * 8 + 4);
          } else {
            v0 = __ldcg(reinterpret_cast<const float4*>(rp + j * 8)); v1 = __ldcg(reinterpret_cast<const float4*>(rp + j * 8 + 4));
            const float4 e0 = ld_nc_f4(e + j * 8), e1 = ld_nc_f4(e + j * 8 + 4);
            v0 = make_float4(v0.x - e0.x, v0.y - e0.y, v0.z - e0.z, v0.w - e0.w);
            v1 = make_float4(v1.x - e1.x, v1.y - e1.y, v1.z - e1.z, v1.w - e1.w);
            __stcg(reinterpret_cast<float4*>(rp + j * 8), v0); __stcg(reinterpret_cast<float4*>(rp + j * 8 + 4), v1);
          }
        }
        const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) part = fmaf(v[i], v[i], part);
        uint4 hv, lv;
        f1::split8(v, hv, lv, mx2);
        const uint32_t a = rrow + (uint32_t)((j ^ (uf & 7)) << 4);
        f1::sts128u(a, hv);
        f1::sts128u(a + kFrames * 128, lv);
      }
      part += __shfl_xor_sync(0xffffffffu, part, 1);
      part += __shfl_xor_sync(0xffffffffu, part, 2);
      if (uq == 0) xn[uf] = part;
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      epi_sync();                                   // operand + xn complete for all frames
      if (lane == 0) tc::mbar_arrive(r_ready);

      // ---- 16 code blocks: running (min, lowest index) for this thread's code lane ------------------------------------
      // The reference takes the argmin of sqrt(d2) (torch.cdist), lowest index among EQUAL ROUNDED roots. A candidate that is
      // smaller by more than 2^-21 relative has a strictly smaller rounded root for certain (the roots differ by 2^-22
      // relative, two roundings move them by at most 2^-24 each), so the running minimum is kept on d2 itself and the square
      // roots are only computed when a candidate falls inside that band below the current minimum -- never, in practice
      // (exact duplicates are equal, not inside the band) -- and once per frame at the end. 9 instructions per candidate
      // instead of 16: the block loop, not the codebook stream, bounded this kernel (ncu: 25 B/clk/SM of stream, issue-bound
      // epilogue warps).
      float bd[32];                                 // d2 of the running minimum
      int bi[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) { bd[i] = INFINITY; bi[i] = 0; }
      const long long cofs = (long long)stage * kCodebookSize + quarter * 32 + lane;
      const uint32_t xn_s = tc::smem_u32(xn + fh * 32);
      for (int blk = 0; blk < kBlocks; ++blk, ++bc) {
        const uint32_t buf = bc & 1u;
        const float e2 = __ldg(p.enorm + cofs + blk * kCodesPerBlock);
        const float ms = __ldg(p.m2s + cofs + blk * kCodesPerBlock);
        const int code = blk * kCodesPerBlock + quarter * 32 + lane;
        tc::mbar_wait(&acc_full[buf], (bc >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t rm[32], rs[32];
        rvqtc::tmem_ld32(tmem_base + lane_off + buf * 128 + fh * 32, rm);
        rvqtc::tmem_ld32(tmem_base + lane_off + buf * 128 + 64 + fh * 32, rs);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
        bool band = false;
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 x4 = lds128(xn_s + (uint32_t)i4 * 16u);
          const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int i = i4 * 4 + j;
            const float dot = fmaf(__uint_as_float(rs[i]), 1.0f / kF16LoScale, __uint_as_float(rm[i]));   // scaled by 2^s
            const float d2 = fmaxf(fmaf(ms, dot, xv[j] + e2), 0.f);
            if (d2 * kBandUp < bd[i]) { bd[i] = d2; bi[i] = code; }   // clearly smaller (codes ascend per thread: ties keep the lowest)
            band |= d2 < bd[i];                                        // smaller, but within the band: decide on the roots below
          }
        }
        if (__builtin_expect(band, 0)) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {     // (unrolled: a dynamic index would move the register arrays to local memory)
            const float dot = fmaf(__uint_as_float(rs[i]), 1.0f / kF16LoScale, __uint_as_float(rm[i]));
            const float d2 = fmaxf(fmaf(ms, dot, xn[fh * 32 + i] + e2), 0.f);
            if (d2 < bd[i] && sqrtf(d2) < sqrtf(bd[i])) { bd[i] = d2; bi[i] = code; }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) bd[i] = sqrtf(bd[i]);   // the cross-lane argmin below compares rounded roots, like the reference
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
    if (uf < nf && (__hisinf(__low2half(mx2)) || __hisinf(__high2half(mx2)))) g_f16_overflow = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256));
  }
}

}  // namespace rvq16
}  // namespace mimi
