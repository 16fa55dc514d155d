// Fused SEANet front end at 24 kHz: waveform -> L0 (conv 1->64, k7) -> ELU -> R1a (64->32, k3) -> ELU ->
// R1b (32->64, k1) + skip -> ELU -> TF32 hi/lo split, written once as the operand of the first strided conv
// (MimiEncoder layers 0, 1.block.1, 1.block.3; modeling_mimi.py:412-451,454-496). The 64-channel 24 kHz
// activations (6.1 MB per audio-second in fp32) never touch HBM: per audio-second the kernel reads 96 KB
// and writes the 12.3 MB split operand.
//
//   One persistent CTA per SM, 16 warps = two groups of 8 warps (256 threads) that ping-pong over time tiles. Inside
//   a group thread (m, ch) owns tile row m and one half of the channels in every phase, so the raw L0 output needed
//   by the skip stays in its registers, and every SM sub-partition has four warps to hide TMEM / barrier / global
//   latencies behind (the first version had one thread per row and two warps per sub-partition: 25 % issue
//   utilisation). Thread 0 of each group issues that group's tcgen05.mma after a 256-thread named barrier.
//     front  L0 on CUDA cores (weights as kernel-parameter constants -> FFMA with constant-bank operands,
//            no loads), ELU, hi/lo split, stored straight into the SWIZZLE_128B K-major operand layout;
//     R1a    3 taps x 2 channel panels = 6 k-blocks; tap tau reads the SAME staged rows through a descriptor
//            whose start address is shifted by tau 128-byte rows (im2col never exists, not even in smem);
//     mid    TMEM -> bias, ELU, split -> operand of R1b (aliases the dead R1a operand);
//     R1b    one k-block;
//     final  TMEM -> bias + skip -> ELU -> split -> 32-column transposes through this warp's own (dead)
//            operand rows -> row-contiguous float4 stores.
//   A tile stages 128 rows (times t0-2 .. t0+125) and keeps the 126 outputs t0 .. t0+125: the two leading
//   rows are the causal halo of the k=3 conv, recomputed instead of carried, so tiles are independent.
//   Arithmetic is 3xTF32 as in tc_gemm2.cuh; K is at most 192 so a single accumulation chunk is used.
#pragma once
#include "tc_gemm2.cuh"

namespace mimi {
namespace f0 {

constexpr int kAdv = 126;                        // outputs kept per tile
constexpr int kPanelRows = 136;
constexpr int kPanelBytes = kPanelRows * 128;    // 17408 = 17 KB, keeps every panel 1024-byte aligned
constexpr int kTileBuf = 4 * kPanelBytes;        // hi panel 0 | hi panel 1 | lo panel 0 | lo panel 1
constexpr int kW1Block = 32 * 128;               // one k-block of W1 (32 output channels x 32 floats), hi or lo
constexpr int kW1Bytes = 6 * kW1Block;           // per hi / lo; stored stacked per k-block: [hi kb | lo kb]
constexpr int kW2Bytes = 64 * 128;               // per hi / lo; stored stacked: [hi | lo]
constexpr int kThreads = 512;
constexpr int kSmem = 1024 + 2 * kTileBuf + 2 * kW1Bytes + 2 * kW2Bytes + 256;
constexpr int kTmemCols = 512;                   // 2 warpgroups x (32 + 32 + 64 + 64) columns

struct Consts {
  float w0[64 * 7];     // L0 weight [64][7]
  float b0[64];
  float b1[32];         // R1a bias
  float b2[64];         // R1b bias
};

struct Params {
  const float* x;            // [B][x_stride] waveform
  long long x_stride;
  const int* len;            // device [B] valid samples per item or nullptr -> uniform_len
  int uniform_len;
  int B;
  int mt_max;                // tiles per item at the longest item
  float* out_hi;             // split output, channels-last rows of 64, halo rows in front
  float* out_lo;
  long long split_item_stride;
  int split_front;
  int lo_bf16;               // 1: out_lo is a bf16 array (mode 7); 3: out_hi and out_lo are fp16 arrays (mode 9, split_f16)
};

__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

__global__ void __launch_bounds__(kThreads, 1)
front_fused_kernel(const __grid_constant__ CUtensorMap tmW1_hi, const __grid_constant__ CUtensorMap tmW1_lo,
                   const __grid_constant__ CUtensorMap tmW2_hi, const __grid_constant__ CUtensorMap tmW2_lo,
                   const __grid_constant__ Consts cst, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // B operands are stored "stacked": k-block kb of W1 is [hi (32 rows) | lo (32 rows)] so that ONE N=64 MMA with
  // A_hi produces the main term (columns 0..31) and the A_hi*W_lo cross term (columns 32..63) together
  uint8_t* w1 = smem + 2 * kTileBuf;               // 6 x (hi 4 KB | lo 4 KB)
  uint8_t* w2 = w1 + 2 * kW1Bytes;                 // hi 8 KB | lo 8 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(w2 + 2 * kW2Bytes);
  uint64_t* acc1_full = bars;           // [2] R1a accumulators of warpgroup g complete (tcgen05.commit)
  uint64_t* acc2_full = bars + 2;       // [2] R1b accumulators complete
  uint64_t* w_ready = bars + 4;         // weights landed (TMA)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int vtiles = p.mt_max * p.B;

  if (threadIdx.x == 0) {
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&acc1_full[g], 1);
      tc::mbar_init(&acc2_full[g], 1);
    }
    tc::mbar_init(w_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // rows 0,1 and 130..135 of every panel are never written by the front phase; they only feed discarded
  // output rows, but keep them finite
  for (int i = threadIdx.x; i < 2 * 4 * 8 * 32; i += kThreads) {
    const int buf = i / (4 * 8 * 32), rem = i % (4 * 8 * 32);
    const int panel = rem / (8 * 32), r8 = (rem / 32) % 8, col = rem % 32;
    const int row = r8 < 2 ? r8 : 128 + r8;
    reinterpret_cast<float*>(smem + buf * kTileBuf + panel * kPanelBytes + row * 128)[col] = 0.f;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  fence_async_smem();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;

  auto item_len = [&](int b) { return p.len ? __ldg(p.len + b) : p.uniform_len; };

  if (threadIdx.x == 0) {
    // resident weights: W1 [32][192] as 6 k-blocks, W2 [64][32]
    tc::prefetch_tmap(&tmW1_hi); tc::prefetch_tmap(&tmW1_lo); tc::prefetch_tmap(&tmW2_hi); tc::prefetch_tmap(&tmW2_lo);
    tc::mbar_expect_tx(w_ready, 2 * kW1Bytes + 2 * kW2Bytes);
    for (int kb = 0; kb < 6; ++kb) {
      tc::tma_load_2d(w1 + (2 * kb) * kW1Block, &tmW1_hi, w_ready, kb * 32, 0);
      tc::tma_load_2d(w1 + (2 * kb + 1) * kW1Block, &tmW1_lo, w_ready, kb * 32, 0);
    }
    tc::tma_load_2d(w2, &tmW2_hi, w_ready, 0, 0);
    tc::tma_load_2d(w2 + kW2Bytes, &tmW2_lo, w_ready, 0, 0);
  }
  {
    // ---- compute warpgroups: 8 warps each; thread = (tile row m, channel half ch) --------------------------------
    const int g = warp >> 3;                       // warpgroup
    const int wq = warp & 3;                       // TMEM lane quarter
    const int ch = (warp >> 2) & 1;                // channel half: L0 / R1b channels [32ch, +32), R1a channels [16ch, +16)
    const int m = wq * 32 + lane;                  // tile row owned by this thread
    const bool issue_warp = (wq == 0) && (ch == 0);   // one elected lane of this warp issues the group's MMAs
    constexpr uint32_t idesc32 = tc::make_idesc(128, 32);
    constexpr uint32_t idesc64 = tc::make_idesc(128, 64);
    constexpr uint32_t idesc128 = tc::make_idesc(128, 128);
    auto wg_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); };
    if (issue_warp) tc::mbar_wait(w_ready, 0);
    uint8_t* buf = smem + g * kTileBuf;
    const uint32_t bufa = tc::smem_u32(buf);
    const uint32_t lane_off = (uint32_t)(wq * 32) << 16;
    const uint32_t tm1 = tmem_base + lane_off + g * 192;           // acc1 main | small (32 + 32 columns)
    const uint32_t tm2 = tm1 + 64;                                  // acc2 main | small (64 + 64 columns)
    const uint32_t stg = bufa + (2 + ch) * kPanelBytes + (wq * 32 + 2) * 128;   // this warp's own rows of lo panel ch
    uint32_t it = 0;
    for (int id = blockIdx.x + g * gridDim.x; id < vtiles; id += 2 * gridDim.x) {
      const int b = id % p.B;
      const int t0 = (id / p.B) * kAdv;
      const int L = item_len(b);
      if (t0 >= L) continue;
      const int t = t0 - 2 + m;                    // time of this thread's row
      // ---- front: L0 (this thread's 32 channels) + ELU + split -> operand row m + 2 of panel ch ------------------
      float a0[32];
      {
        const float* xb = p.x + (long long)b * p.x_stride;
        float xv[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const int tt = t - 6 + k;
          xv[k] = (tt >= 0 && tt < L) ? __ldg(xb + tt) : 0.f;
        }
        if (ch == 0) {   // pull the next tile's samples towards L1 while this tile computes
          const int nid = id + 2 * gridDim.x;
          if (nid < vtiles) {
            const int nt = (nid / p.B) * kAdv - 2 + m;
            if (nt >= 0 && nt < p.uniform_len)
              asm volatile("prefetch.global.L1 [%0];" ::"l"(p.x + (long long)(nid % p.B) * p.x_stride + nt));
          }
        }
        // two copies so that every weight is a compile-time constant-bank operand of its FFMA
        if (ch == 0) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float s0 = cst.b0[c];
#pragma unroll
            for (int k = 0; k < 7; ++k) s0 = fmaf(cst.w0[c * 7 + k], xv[k], s0);
            a0[c] = s0;
          }
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float s0 = cst.b0[32 + c];
#pragma unroll
            for (int k = 0; k < 7; ++k) s0 = fmaf(cst.w0[(32 + c) * 7 + k], xv[k], s0);
            a0[c] = s0;
          }
        }
        const int row = m + 2;
        const int key = row & 7;
        const bool live = t >= 0;                  // rows before the item start are the conv's zero padding
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 h4, l4;
          split_tf32(live ? elu_fast(a0[q * 4 + 0]) : 0.f, h4.x, l4.x);
          split_tf32(live ? elu_fast(a0[q * 4 + 1]) : 0.f, h4.y, l4.y);
          split_tf32(live ? elu_fast(a0[q * 4 + 2]) : 0.f, h4.z, l4.z);
          split_tf32(live ? elu_fast(a0[q * 4 + 3]) : 0.f, h4.w, l4.w);
          const uint32_t off = bufa + (uint32_t)(row * 128 + ((q ^ key) << 4));
          sts128(off + ch * kPanelBytes, h4);
          sts128(off + (2 + ch) * kPanelBytes, l4);
        }
        fence_async_smem();
        wg_sync();
        if (issue_warp && tc::elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t main1 = tmem_base + g * 192, small1 = main1 + 32;
          const uint32_t d_buf = tc::desc_lo(bufa), d_w1 = tc::desc_lo(tc::smem_u32(w1));
#pragma unroll
          for (int kb = 0; kb < 6; ++kb) {
            const int tau = kb >> 1, pn = kb & 1;
            const uint32_t a_hi = d_buf + ((pn * kPanelBytes + tau * 128) >> 4);
            const uint32_t a_lo = a_hi + ((2 * kPanelBytes) >> 4);
            const uint32_t b_st = d_w1 + (((2 * kb) * kW1Block) >> 4);        // [hi | lo] stacked, N = 64
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              // columns [0,32) += A_hi W_hi^T (main), columns [32,64) += A_hi W_lo^T; then [32,64) += A_lo W_hi^T
              tc::umma_tf32_lo(main1, a_hi + 2 * k, b_st + 2 * k, idesc64, (uint32_t)((kb | k) != 0));
              tc::umma_tf32_lo(small1, a_lo + 2 * k, b_st + 2 * k, idesc32, 1u);
            }
          }
          tc::umma_commit(&acc1_full[g]);
        }
        __syncwarp();
      }
      // ---- mid: R1a accumulators (this thread's 16 channels) -> bias, ELU, split -> R1b operand row m -----------------
      {
        tc::mbar_wait(&acc1_full[g], it & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t rm[16], rs[16];
        tc2::tmem_ld16_nowait(tm1 + ch * 16, rm);
        tc2::tmem_ld16_nowait(tm1 + 32 + ch * 16, rs);
        tc2::tmem_ld_wait();
        const int key = m & 7;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 h4, l4;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float bias = ch == 0 ? cst.b1[q * 4 + j] : cst.b1[16 + q * 4 + j];
            v[j] = elu_fast(__uint_as_float(rm[q * 4 + j]) + __uint_as_float(rs[q * 4 + j]) + bias);
          }
          split_tf32(v[0], h4.x, l4.x); split_tf32(v[1], h4.y, l4.y); split_tf32(v[2], h4.z, l4.z); split_tf32(v[3], h4.w, l4.w);
          const uint32_t off = bufa + (uint32_t)(m * 128 + (((ch * 4 + q) ^ key) << 4));
          sts128(off, h4);
          sts128(off + kPanelBytes, l4);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        fence_async_smem();
        wg_sync();
        if (issue_warp && tc::elect_one()) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t a_hi = tc::desc_lo(bufa);                       // R1 operand aliases hi panel 0 / hi panel 1
          const uint32_t a_lo = a_hi + (kPanelBytes >> 4);
          const uint32_t b_st = tc::desc_lo(tc::smem_u32(w2));           // [hi | lo] stacked, N = 128
          const uint32_t main2 = tmem_base + g * 192 + 64, small2 = main2 + 64;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            tc::umma_tf32_lo(main2, a_hi + 2 * k, b_st + 2 * k, idesc128, (uint32_t)(k != 0));
            tc::umma_tf32_lo(small2, a_lo + 2 * k, b_st + 2 * k, idesc64, 1u);
          }
          tc::umma_commit(&acc2_full[g]);
        }
        __syncwarp();
      }
      // ---- final: R1b accumulators (this thread's 32 channels) + bias + skip -> ELU -> split -> coalesced stores ----
      {
        tc::mbar_wait(&acc2_full[g], it & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const long long obase = (long long)b * p.split_item_stride + (long long)p.split_front * 64 + ch * 32;
        uint32_t rm[32], rs[32];
        tmem_ld32(tm2 + ch * 32, rm);
        tmem_ld32(tm2 + 64 + ch * 32, rs);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        const int key = lane & 7;
        float4 lq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 h4;
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int c = q * 4 + j;
            const float bias = ch == 0 ? cst.b2[c] : cst.b2[32 + c];
            const float pre = (__uint_as_float(rm[c]) + __uint_as_float(rs[c])) + bias + a0[c];
            v[j] = elu_fast(pre);
          }
          if (p.lo_bf16 == 3) {
            split_f16(v[0], h4.x, lq[q].x); split_f16(v[1], h4.y, lq[q].y); split_f16(v[2], h4.z, lq[q].z); split_f16(v[3], h4.w, lq[q].w);
          } else {
            split_tf32(v[0], h4.x, lq[q].x); split_tf32(v[1], h4.y, lq[q].y); split_tf32(v[2], h4.z, lq[q].z); split_tf32(v[3], h4.w, lq[q].w);
          }
          sts128(stg + (uint32_t)(lane * 8 + (q ^ key)) * 16u, h4);
        }
        // hi piece, then lo piece, through the same 4 KB of this warp's own (dead) operand rows
        for (int pass = 0; pass < 2; ++pass) {
          __syncwarp();
          float4 tv[8];
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = i8 * 4 + (lane >> 3);        // row inside this warp's 32
            tv[i8] = lds128(stg + (uint32_t)(r * 8 + ((lane & 7) ^ (r & 7))) * 16u);
          }
          float* outp = pass == 0 ? p.out_hi : p.out_lo;
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int r = i8 * 4 + (lane >> 3);
            const int mm = wq * 32 + r;
            const int tt = t0 - 2 + mm;
            if (mm >= 2 && tt < L) {
              const long long o = obase + (long long)tt * 64 + (lane & 7) * 4;
              if (p.lo_bf16 == 3)
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(outp) + o) =
                    make_uint2(pack_f16x2(tv[i8].x, tv[i8].y), pack_f16x2(tv[i8].z, tv[i8].w));
              else if (pass == 1 && p.lo_bf16)
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(outp) + o) =
                    make_uint2(pack_bf16x2(tv[i8].x, tv[i8].y), pack_bf16x2(tv[i8].z, tv[i8].w));
              else
                *reinterpret_cast<float4*>(outp + o) = tv[i8];
            }
          }
          __syncwarp();
          if (pass == 0) {
#pragma unroll
            for (int q = 0; q < 8; ++q) sts128(stg + (uint32_t)(lane * 8 + (q ^ key)) * 16u, lq[q]);
          }
        }
      }
      ++it;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

}  // namespace f0
}  // namespace mimi
