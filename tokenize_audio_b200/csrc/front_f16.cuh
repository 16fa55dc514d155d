// Fused SEANet front end at 24 kHz, fp16-pair generation (mode 9): waveform -> L0 (conv 1->64, k7) -> ELU -> R1a (64->32, k3)
// -> ELU -> R1b (32->64, k1) + skip -> ELU -> fp16 hi/lo split, written once as the operand of the first strided conv
// (MimiEncoder layers 0, 1.block.1, 1.block.3; modeling_mimi.py:412-451,454-496). Same structure as front_fused.cuh (one
// persistent CTA per SM, two groups of 8 warps ping-ponging over 128-row time tiles of which 126 are kept, thread = (tile row,
// channel half), L0 on CUDA cores, R1a / R1b on tcgen05 with the three taps of R1a as row-shifted descriptors over the same
// staged rows), rebuilt around the instruction count, which is what bounded the first version (ncu: issue slots 55 % busy with
// 72 instructions per output element, tensor pipe 22 %):
//   * the internal operands are fp16 pairs like every other GEMM operand of mode 9 (common.cuh: split_f16) -- 64 channels = one
//     128-byte SWIZZLE_128B row, so a tile is 2 panels instead of 4 and a thread stores 8 + 4 instead of 16 + 8 16-byte pieces;
//     hi*W_hi + hi*W_lo + lo*W_hs run as three kind::f16 MMAs per K = 16 step into ONE accumulator (half the TMEM loads, no
//     main + cross-term add), weights row-scaled by a power of two and unscaled by the per-column affine as in tc_gemm5.cuh;
//   * the final phase packs hi and lo to fp16 once, stages BOTH in one pass through a per-warp 4 KB tile and stores 16 bytes
//     per lane (the first version staged fp32 twice and converted again on the way out);
//   * one range check per thread at kernel end (a running half2 maximum) instead of a compare + predicated flag store per value;
//   * the weights arrive as a ready-made shared-memory image (swizzle applied on the host): no tensor maps, no TMA barrier;
//   * tiles come from the items' lengths (item-major, only tiles that exist), walked incrementally: no division per tile and
//     no skipped (item, tile) pairs unbalancing the static round-robin; the next tile's samples are loaded under the current one.
#pragma once
#include "front_fused.cuh"
#include "tc_gemm5.cuh"

namespace mimi {
namespace f1 {

constexpr int kAdv = 126;                        // outputs kept per tile
constexpr int kPanelRows = 136;
constexpr int kPanelBytes = kPanelRows * 128;    // 17408, a multiple of 1024
constexpr int kGroupBuf = 2 * kPanelBytes;       // hi panel | lo panel
constexpr int kWBlock = 32 * 128;                // 32 weight rows of 64 halfs
constexpr int kW1Bytes = 9 * kWBlock;            // [tap][hi | lo | hs][32 rows]
constexpr int kW2Part = 64 * 128;                // 64 rows, the first 64 bytes (K = 32) of each used
constexpr int kW2Bytes = 3 * kW2Part;            // hi | lo | hs
constexpr int kWBytes = kW1Bytes + kW2Bytes;     // 61440
constexpr int kStageBytes = 16 * 4096;           // one 32 x 128 B transpose tile per warp
constexpr int kThreads = 512;
constexpr int kSmem = 1024 + 2 * kGroupBuf + kWBytes + kStageBytes + 64;
constexpr int kTmemCols = 256;                   // per group 128: acc1 (32 columns) | acc2 (64 columns)

struct Consts {
  float w0[64 * 7];     // L0 weight [64][7]
  float b0[64];
  float m1[32], a1[32]; // R1a: out = acc * m1 + a1 (weight unscaling, bias)
  float m2[64], a2[64]; // R1b
};

struct Params {
  const float* x;            // [B][x_stride] waveform
  long long x_stride;
  const int* len;            // device [B] valid samples per item or nullptr -> uniform_len
  int uniform_len;
  int B;
  int total_tiles;           // sum over the items of ceil(len / 126)
  const uint4* wimg;         // kWBytes: the shared-memory image of the resident weights
  uint16_t* out_hi;          // split output, channels-last rows of 64 halfs, halo rows in front
  uint16_t* out_lo;
  long long split_item_stride;
  int split_front;
};

__device__ __forceinline__ void sts128u(uint32_t saddr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128u(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}
// kind::f16, cta_group::1, SWIZZLE_128B K-major operands (descriptor high word tc::kDescHi)
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(tc::kDescHi)
      : "memory");
}

// keeps descriptor arithmetic inside the (single-lane) issue branch: without it the compiler computes every descriptor in all
// 512 threads ahead of the branch and spills them (17 STL per tile per thread in the first build)
__device__ __forceinline__ uint32_t pin(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

// eight values -> fp16 pair: hv = fp16(v) packed, lv = fp16((v - hv) * 2048) packed; mx2 tracks max |hv| (range check)
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hv, uint4& lv, __half2& mx2) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h[i] = pack_f16x2(v[2 * i], v[2 * i + 1]);
    const __half2 g = *reinterpret_cast<const __half2*>(&h[i]);
    const float2 f = __half22float2(g);
    l[i] = pack_f16x2((v[2 * i] - f.x) * kF16LoScale, (v[2 * i + 1] - f.y) * kF16LoScale);
    mx2 = __hmax2(mx2, __habs2(g));
  }
  hv = make_uint4(h[0], h[1], h[2], h[3]);
  lv = make_uint4(l[0], l[1], l[2], l[3]);
}

// Front phase of one thread: L0 for its 32 channels (every weight a compile-time constant-bank operand of its FFMA), ELU, fp16
// split, operand row stores -- eight channels at a time -- and the skip as the final phase wants it (L0 output + R1b bias: the
// add costs nothing extra here and saves an add and a constant load there). `rowa`: shared address of the thread's row in the
// hi panel; `live` false: the row lies before the item start, i.e. it is the conv's zero padding.
template <int CH>
__device__ __forceinline__ void front_phase(const Consts& cst, const float (&xv)[7], float (&skip)[32], bool live, uint32_t rowa,
                                            int key, __half2& mx2) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = CH * 32 + q * 8 + j;
      float s0 = cst.b0[c];
#pragma unroll
      for (int k = 0; k < 7; ++k) s0 = fmaf(cst.w0[c * 7 + k], xv[k], s0);
      v[j] = elu_fast(s0);
      skip[q * 8 + j] = s0 + cst.a2[c];
    }
    uint4 hv, lv;
    split8(v, hv, lv, mx2);
    const uint32_t off = rowa + (uint32_t)(((CH * 4 + q) ^ key) << 4);
    if (live) {
      sts128u(off, hv);
      sts128u(off + kPanelBytes, lv);
    } else {
      sts128u(off, make_uint4(0, 0, 0, 0));
      sts128u(off + kPanelBytes, make_uint4(0, 0, 0, 0));
    }
  }
}

// rotation of the low three row bits: the swizzle key of the transpose tile (rows r, r + 1 land in opposite halves of a row)
__device__ __forceinline__ int stage_key(int r) { return ((r & 1) << 2) | ((r >> 1) & 3); }

__global__ void __launch_bounds__(kThreads, 1)
front_f16_kernel(const __grid_constant__ Consts cst, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem + 2 * kGroupBuf;             // W1: 9 blocks of 4 KB; W2: 3 parts of 8 KB
  uint8_t* stage = wsm + kWBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + kStageBytes);
  uint64_t* acc1_full = bars;           // [2] R1a accumulators of group g complete (tcgen05.commit)
  uint64_t* acc2_full = bars + 2;       // [2] R1b accumulators complete
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&acc1_full[g], 1);
      tc::mbar_init(&acc2_full[g], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // resident weights: a straight copy of the host-made image
  for (int i = threadIdx.x; i < kWBytes / 16; i += kThreads)
    *reinterpret_cast<uint4*>(wsm + i * 16) = __ldg(p.wimg + i);
  // rows 0, 1 of the panels are never written by the front phase and only feed discarded output rows; keep them finite
  if (threadIdx.x < 2 * 2 * 2 * 8) {
    const int i = threadIdx.x;
    const int buf = i >> 5, panel = (i >> 4) & 1, row = (i >> 3) & 1, chunk = i & 7;
    *reinterpret_cast<uint4*>(smem + buf * kGroupBuf + panel * kPanelBytes + row * 128 + chunk * 16) = make_uint4(0, 0, 0, 0);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;

  const int g = warp >> 3;                       // group
  const int wq = warp & 3;                       // TMEM lane quarter
  const int ch = (warp >> 2) & 1;                // channel half: L0 / R1b channels [32 ch, +32), R1a channels [16 ch, +16)
  const int m = wq * 32 + lane;                  // tile row owned by this thread
  const bool issue_warp = (wq == 0) && (ch == 0);   // one elected lane of this warp issues the group's MMAs
  constexpr uint32_t idesc32 = tcp::make_idesc_f16(128, 32);
  constexpr uint32_t idesc64 = tcp::make_idesc_f16(128, 64);
  auto wg_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(1 + g) : "memory"); };
  const uint32_t bufa = tc::smem_u32(smem + g * kGroupBuf);             // hi panel; lo panel at + kPanelBytes
  const uint32_t tm1 = tmem_base + ((uint32_t)(wq * 32) << 16) + g * 128;     // acc1: 32 columns
  const uint32_t tm2 = tm1 + 32;                                               // acc2: 64 columns
  const uint32_t stg = tc::smem_u32(stage) + (uint32_t)warp * 4096u;
  const uint32_t d_hi = tc::desc_lo(bufa), d_lo = tc::desc_lo(bufa + kPanelBytes);
  const uint32_t d_w1 = tc::desc_lo(tc::smem_u32(wsm)), d_w2 = tc::desc_lo(tc::smem_u32(wsm + kW1Bytes));
  __half2 mx2 = __floats2half2_rn(0.f, 0.f);     // running max |hi| of everything this thread splits

  // ---- tile walk: linear tile id -> (item b, tile of the item), items in order, only tiles that exist ---------------------
  auto item_len = [&](int b) { return p.len ? __ldg(p.len + b) : p.uniform_len; };
  const int step = 2 * (int)gridDim.x;
  int id = (int)blockIdx.x + g * (int)gridDim.x;
  int b = 0, base = 0, L = item_len(0), cnt = (L + kAdv - 1) / kAdv;   // item of the tile in `xv`, its first tile id, samples, tiles
  float xv[7];
  auto seek_and_load = [&]() {                   // id < total_tiles
    while (id >= base + cnt) {
      base += cnt;
      ++b;
      L = item_len(b);
      cnt = (L + kAdv - 1) / kAdv;
    }
    const int t = (id - base) * kAdv - 2 + m;
    const float* xp = p.x + (long long)b * p.x_stride + (t - 6);
#pragma unroll
    for (int k = 0; k < 7; ++k) xv[k] = ((unsigned)(t - 6 + k) < (unsigned)L) ? __ldg(xp + k) : 0.f;
  };
  if (id < p.total_tiles) seek_and_load();

  uint32_t it = 0;
  while (id < p.total_tiles) {
    const int cb = b, cL = L, t0 = (id - base) * kAdv;     // the tile being computed
    const int t = t0 - 2 + m;                              // time of this thread's row
    // ---- front: L0 (this thread's 32 channels) + ELU + split -> operand row m + 2 ---------------------------------------------
    float a0[32];                                          // L0 output + R1b bias: the skip of the final phase
    {
      const int row = m + 2;
      if (ch == 0) front_phase<0>(cst, xv, a0, t >= 0, bufa + (uint32_t)(row * 128), row & 7, mx2);
      else front_phase<1>(cst, xv, a0, t >= 0, bufa + (uint32_t)(row * 128), row & 7, mx2);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      wg_sync();
      if (issue_warp && tc::elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc1 = pin(tmem_base + g * 128);
        const uint32_t p_hi = pin(d_hi), p_lo = pin(d_lo), p_w1 = pin(d_w1);
#pragma unroll
        for (int tau = 0; tau < 3; ++tau) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            // tap tau = the same staged rows, start address tau rows further; K = 16 halfs = 32 bytes per step
            const uint32_t a_hi = p_hi + (uint32_t)(tau * 8 + 2 * k), a_lo = p_lo + (uint32_t)(tau * 8 + 2 * k);
            const uint32_t w = p_w1 + (uint32_t)((tau * 3 * kWBlock) >> 4) + 2 * k;
            umma_f16(acc1, a_hi, w, idesc32, (uint32_t)((tau | k) != 0));
            umma_f16(acc1, a_hi, w + (kWBlock >> 4), idesc32, 1u);
            umma_f16(acc1, a_lo, w + ((2 * kWBlock) >> 4), idesc32, 1u);
          }
        }
        tc::umma_commit(&acc1_full[g]);
      }
      __syncwarp();
    }
    // the next tile's samples travel while this tile waits for its MMAs (issued after the proxy fence above, which would
    // otherwise wait for them)
    id += step;
    if (id < p.total_tiles) seek_and_load();
    // ---- mid: R1a accumulators (this thread's 16 channels) -> affine, ELU, split -> R1b operand row m (aliases the panels) ----
    {
      tc::mbar_wait(&acc1_full[g], it & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[16];
      tc2::tmem_ld16_nowait(tm1 + ch * 16, r);
      tc2::tmem_ld_wait();
      const int key = m & 7;
      const uint32_t rowa = bufa + (uint32_t)(m * 128);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = ch * 16 + q * 8 + j;
          v[j] = elu_fast(fmaf(__uint_as_float(r[q * 8 + j]), cst.m1[c], cst.a1[c]));
        }
        uint4 hv, lv;
        split8(v, hv, lv, mx2);
        const uint32_t off = rowa + (uint32_t)(((ch * 2 + q) ^ key) << 4);
        sts128u(off, hv);
        sts128u(off + kPanelBytes, lv);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      wg_sync();
      if (issue_warp && tc::elect_one()) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc2 = pin(tmem_base + g * 128 + 32);
        const uint32_t p_hi = pin(d_hi), p_lo = pin(d_lo), p_w2 = pin(d_w2);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          umma_f16(acc2, p_hi + 2 * k, p_w2 + 2 * k, idesc64, (uint32_t)(k != 0));
          umma_f16(acc2, p_hi + 2 * k, p_w2 + (kW2Part >> 4) + 2 * k, idesc64, 1u);
          umma_f16(acc2, p_lo + 2 * k, p_w2 + ((2 * kW2Part) >> 4) + 2 * k, idesc64, 1u);
        }
        tc::umma_commit(&acc2_full[g]);
      }
      __syncwarp();
    }
    // ---- final: R1b accumulators (this thread's 32 channels) -> affine + skip -> ELU -> split -> transposed 16-byte stores ----
    {
      tc::mbar_wait(&acc2_full[g], it & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t r[32];
      f0::tmem_ld32(tm2 + ch * 32, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      const int skey = stage_key(lane);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = q * 8 + j;
          v[j] = elu_fast(fmaf(__uint_as_float(r[c]), cst.m2[ch * 32 + c], a0[c]));
        }
        uint4 hv, lv;
        split8(v, hv, lv, mx2);
        sts128u(stg + (uint32_t)(lane * 8 + (q ^ skey)) * 16u, hv);
        sts128u(stg + (uint32_t)(lane * 8 + ((4 + q) ^ skey)) * 16u, lv);
      }
      __syncwarp();
      const long long obase = (long long)cb * p.split_item_stride + (long long)p.split_front * 64 + ch * 32 + (lane & 3) * 8;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rr = i * 8 + (lane >> 2);            // row inside this warp's 32
        const int rkey = stage_key(rr);
        const uint4 hv = lds128u(stg + (uint32_t)(rr * 8 + ((lane & 3) ^ rkey)) * 16u);
        const uint4 lv = lds128u(stg + (uint32_t)(rr * 8 + ((4 + (lane & 3)) ^ rkey)) * 16u);
        const int mm = wq * 32 + rr;
        const int tt = t0 - 2 + mm;
        if (mm >= 2 && tt < cL) {
          const long long o = obase + (long long)tt * 64;
          __stcs(reinterpret_cast<uint4*>(p.out_hi + o), hv);
          __stcs(reinterpret_cast<uint4*>(p.out_lo + o), lv);
        }
      }
      __syncwarp();
    }
    ++it;
  }
  // |hi| is inf exactly when a value was beyond fp16's range; rows m < 2 only ever hold discarded halo outputs
  if (m >= 2 && (__hisinf(__low2half(mx2)) || __hisinf(__high2half(mx2)))) g_f16_overflow = 1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

}  // namespace f1
}  // namespace mimi
