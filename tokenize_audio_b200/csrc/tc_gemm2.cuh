// Pieces shared by the tcgen05 GEMM kernels (tc_gemm5.cuh), the fused front end and the tensor-core attention: the tile
// schedule (compact tile lists, k-block visiting order), chunk drains TMEM -> registers, and the tile finish (bias / GELU /
// LayerScale per thread = per row, a 16-column transpose through a swizzled shared-memory tile, row-contiguous float4 residual
// loads and raw / split stores). Plus a one-tile hardware probe for the SWIZZLE_128B row-shift property the front end uses.
#pragma once
#include "tc_gemm.cuh"

namespace mimi {
namespace tc2 {

using tc::Epilogue;
using tc::kBK;
using tc::kBM;
using tc::kChunkKB;
using tc::kUmmaK;

constexpr int kSmemMax = 232448;                    // 227 KB opt-in limit per CTA

struct Sched {
  int B;            // items
  int mt_max;       // m-tiles per item at the longest item
  int ntn;          // n-tiles = N / BN
  // k-block visiting order of a conv with k = G * s taps over cp = C_in / 32 channel panels (G = 0: linear order).
  // Taps tau and tau + s of neighbouring output rows read the SAME input rows; visiting them back to back makes the
  // second fetch an L2 hit. In linear order they are s * cp k-blocks (x 148 CTAs x 64 KB: more than the L2) apart and
  // the activations are read from DRAM G times (ncu: 2.4 GB instead of 1.2 GB for D2).
  int G, s, cp;
  // Compact list of the 128-row tiles that exist (ragged batches): entry = item << 20 | m-tile, m-tile major. Without it
  // (nullptr) the tiles are the mt_max x B grid and the ones past an item's end are skipped -- which leaves the CTAs of a
  // static round-robin with unequal numbers of real tiles (ncu: SMs idle 20-30 % of a launch on a ragged batch).
  const int* tiles;
  int ntiles;
};

__device__ __forceinline__ int sched_tiles(const Sched& sc) { return sc.tiles ? sc.ntiles : sc.mt_max * sc.B; }
// 128-row tile t -> (item, first row, rows of the item); false when the tile lies past the item's end
__device__ __forceinline__ bool sched_tile(const Sched& sc, const Epilogue& ep, int t, int& b, int& m0, int& Lout) {
  if (sc.tiles) {
    const int e = __ldg(sc.tiles + t);
    b = e >> 20;
    m0 = (e & 0xFFFFF) * kBM;
  } else {
    b = t % sc.B;
    m0 = (t / sc.B) * kBM;
  }
  const int Lin = ep.len_in ? __ldg(ep.len_in + b) : ep.uniform_len_in;
  Lout = (Lin + ep.conv_stride - 1) / ep.conv_stride;
  return m0 < Lout;
}

// i-th k-block to visit -> linear k-block index (tau * cp + channel panel), tau = ph + s * dq with dq fastest, then the
// channel panel, then the tap phase ph. (The GEMM producer walks this order incrementally; this is the closed form.)
__device__ __forceinline__ int kblock_order(const Sched& sc, int i) {
  if (sc.G <= 1) return i;
  const int dq = i % sc.G, t = i / sc.G;
  const int cb = t % sc.cp, ph = t / sc.cp;
  return (ph + sc.s * dq) * sc.cp + cb;
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// acc[0 .. NCOL) += TMEM[taddr .. taddr + NCOL) for this thread's lane; two loads in flight
template <int NCOL>
__device__ __forceinline__ void drain_add(uint32_t taddr, float (&acc)[NCOL]) {
  static_assert(NCOL % 16 == 0, "NCOL");
  if constexpr (NCOL == 16) {
    uint32_t r[16];
    tmem_ld16_nowait(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] += __uint_as_float(r[i]);
  } else {
#pragma unroll
    for (int c0 = 0; c0 < NCOL; c0 += 32) {
      uint32_t r0[16], r1[16];
      tmem_ld16_nowait(taddr + (uint32_t)c0, r0);
      tmem_ld16_nowait(taddr + (uint32_t)c0 + 16u, r1);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[c0 + i] += __uint_as_float(r0[i]);
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[c0 + 16 + i] += __uint_as_float(r1[i]);
    }
  }
}

// Epilogue geometry of one thread: HALF accumulator columns, finished in pieces of PC columns that go through a
// swizzled 32 x PC staging tile of the warp (transpose: thread = row  ->  lane = float4 of a row).
template <int HALF, int PC>
struct EpiGeom {
  static constexpr int NP = HALF / PC;
  static constexpr int LPR = PC / 4;
  static constexpr int RPI = 32 / LPR;
  static constexpr int IT = 32 / RPI;
};

// pull this thread's residual rows towards L2 while the MMAs run (the loads in finish_tile then miss only L1)
template <int HALF, int PC>
__device__ __forceinline__ void prefetch_residual(const Epilogue& ep, int b, int row_base, int ncol0, int Lout, int lane) {
  using G = EpiGeom<HALF, PC>;
  const int rr = lane / G::LPR, cj = lane % G::LPR;
  if (ep.res && cj == 0) {
    const float* rp = ep.res + (long long)b * ep.raw_item_stride + ncol0;
#pragma unroll
    for (int p = 0; p < G::NP; ++p)
#pragma unroll
      for (int it = 0; it < G::IT; ++it) {
        const int row = row_base + it * G::RPI + rr;
        if (row < Lout) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + (long long)row * ep.N + p * PC));
      }
  }
}

// streaming 8-byte store under a predicate (no branch)
__device__ __forceinline__ void st_cs_v2_if(void* p, uint32_t a, uint32_t b, uint32_t pred) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.global.cs.v2.b32 [%0], {%1, %2};\n\t}"
               ::"l"(p), "r"(a), "r"(b), "r"(pred) : "memory");
}

// Finish a tile: thread (lane) holds acc[0, HALF) = columns ncol0 + [0, HALF) of row row_base + lane.
//   phase 1 (thread = row): v = acc * cmul[c] + cadd[c] (the per-column affine the host folds weight unscaling, bias and
//            LayerScale into), GELU(erf) for fc1, into the warp's swizzled staging tile;
//   phase 2 (lane = one float4 of a row, 32 / LPR rows per instruction): + residual, raw store, ELU, split store -- row-contiguous
//            64-byte (fp32) / 32-byte (16-bit) segments.
// Everything that is uniform over the tile (which outputs exist, whether there is a residual / activation) is decided once per
// PC-column piece, outside the unrolled loops, and the row offsets are computed once per tile: the first version re-tested
// every flag and rebuilt every 64-bit address per float4 and spent 43 instructions per output element here (ncu, round 2).
// LOB = 1: split as TF32 hi (fp32) + bf16 lo; LOB = 3: fp16 pair, with the fp16 range check folded into one half2 max per pair.
template <int HALF, int PC, int LOB>
__device__ __forceinline__ void finish_tile(const Epilogue& ep, float (&acc)[HALF], int b, int row_base, int ncol0, int Lout,
                                            uint32_t stg, int lane) {
  using G = EpiGeom<HALF, PC>;
  constexpr int NP = G::NP, LPR = G::LPR, RPI = G::RPI, IT = G::IT;
  const int rr = lane / LPR;                     // coalesced phase: row within an RPI-row group
  const int cj = lane % LPR;                     //                  float4 index inside the PC-wide piece
  // rows of this thread's coalesced phase that exist (bit `it`): inside the tile's item, and -- in a flattened launch -- inside
  // the length of the item the row falls into; and their element offsets row * N + 4 cj
  uint32_t live = 0;
  // addresses = a tile-uniform 64-bit base (item, first row of this warp's 32, first column) + a per-thread 32-bit offset inside
  // the tile: one add per access instead of a 64-bit multiply-add chain (the first layout rebuilt row * N + column in 64 bits
  // for every float4: 3.5 of the 33 instructions per output element of the k = 1 convs)
  uint32_t toff[IT];
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int row = row_base + it * RPI + rr;
    bool ok = row < Lout;
    if (ok && ep.flat_rows > 0 && ep.flat_len) ok = (row % ep.flat_rows) < __ldg(ep.flat_len + row / ep.flat_rows);
    live |= (uint32_t)ok << it;
    toff[it] = (uint32_t)((it * RPI + rr) * ep.N + cj * 4);
  }
  const float* __restrict__ cm = ep.cmul + ncol0;
  const float* __restrict__ ca = ep.cadd + ncol0;
  const long long tile_off = (long long)row_base * ep.N + ncol0;
  const long long raw_base = (long long)b * ep.raw_item_stride + tile_off;
  const long long split_base = (long long)b * ep.split_item_stride + (long long)ep.split_front * ep.N + tile_off;
  const float* resp = ep.res ? ep.res + raw_base : nullptr;
  float* rawp = ep.out_raw ? ep.out_raw + raw_base : nullptr;
  uint16_t* hi16 = reinterpret_cast<uint16_t*>(ep.out_hi) + split_base;
  uint16_t* lo16 = reinterpret_cast<uint16_t*>(ep.out_lo) + split_base;
  const bool do_split = ep.out_hi != nullptr, do_elu = ep.elu_split != 0, do_act = ep.act == 1, do_act_fast = ep.act == 2;
  const int wkey = (LPR == 8) ? (lane & 7) : ((lane >> 1) & (LPR - 1));
  __half2 mx2 = __floats2half2_rn(0.f, 0.f);     // LOB 3: running max |hi| of everything this thread stores (range check)
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float4 resv[IT];
    if (resp) {
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        resv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((live >> it) & 1u) resv[it] = *reinterpret_cast<const float4*>(resp + (toff[it] + p * PC));
      }
    }
    // ---- phase 1: affine (+ GELU) -> staging tile ----
    float4 v[LPR];
#pragma unroll
    for (int j = 0; j < LPR; ++j) {
      const float4 m = ld_nc_f4(cm + p * PC + 4 * j), a = ld_nc_f4(ca + p * PC + 4 * j);
      v[j] = make_float4(fmaf(acc[p * PC + 4 * j], m.x, a.x), fmaf(acc[p * PC + 4 * j + 1], m.y, a.y),
                         fmaf(acc[p * PC + 4 * j + 2], m.z, a.z), fmaf(acc[p * PC + 4 * j + 3], m.w, a.w));
    }
    if (do_act_fast) {
#pragma unroll
      for (int j = 0; j < LPR; ++j) { v[j].x = gelu_fast(v[j].x); v[j].y = gelu_fast(v[j].y); v[j].z = gelu_fast(v[j].z); v[j].w = gelu_fast(v[j].w); }
    } else if (do_act) {
#pragma unroll
      for (int j = 0; j < LPR; ++j) { v[j].x = gelu_erf(v[j].x); v[j].y = gelu_erf(v[j].y); v[j].z = gelu_erf(v[j].z); v[j].w = gelu_erf(v[j].w); }
    }
#pragma unroll
    for (int j = 0; j < LPR; ++j) sts128(stg + (uint32_t)(lane * LPR + (j ^ wkey)) * 16u, v[j]);
    __syncwarp();
    // ---- phase 2: transposed read, residual, raw store, ELU, split store ----
    float4 tv[IT];
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int r = it * RPI + rr;               // row within this warp's 32
      const int rkey = (LPR == 8) ? (r & 7) : ((r >> 1) & (LPR - 1));
      tv[it] = lds128(stg + (uint32_t)(r * LPR + (cj ^ rkey)) * 16u);
    }
    if (resp) {
#pragma unroll
      for (int it = 0; it < IT; ++it) { tv[it].x += resv[it].x; tv[it].y += resv[it].y; tv[it].z += resv[it].z; tv[it].w += resv[it].w; }
    }
    if (rawp) {
#pragma unroll
      for (int it = 0; it < IT; ++it)
        if ((live >> it) & 1u) __stcs(reinterpret_cast<float4*>(rawp + (toff[it] + p * PC)), tv[it]);   // streaming stores: the
        // outputs are re-read only by the NEXT kernel, after more traffic than the L2 holds (D1 -4 %, R2b -6 %; an evict-first
        // residual LOAD was 15 % slower: it fights the L2 prefetch issued at tile start)
    }
    if (do_split) {
      if (do_elu) {
#pragma unroll
        for (int it = 0; it < IT; ++it) { tv[it].x = elu_fast(tv[it].x); tv[it].y = elu_fast(tv[it].y); tv[it].z = elu_fast(tv[it].z); tv[it].w = elu_fast(tv[it].w); }
      }
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const uint32_t o = toff[it] + p * PC;
        if (LOB == 3) {
          const uint32_t h01 = pack_f16x2(tv[it].x, tv[it].y), h23 = pack_f16x2(tv[it].z, tv[it].w);
          const __half2 g01 = *reinterpret_cast<const __half2*>(&h01), g23 = *reinterpret_cast<const __half2*>(&h23);
          const float2 f01 = __half22float2(g01), f23 = __half22float2(g23);
          const uint2 lo2 = make_uint2(pack_f16x2((tv[it].x - f01.x) * kF16LoScale, (tv[it].y - f01.y) * kF16LoScale),
                                       pack_f16x2((tv[it].z - f23.x) * kF16LoScale, (tv[it].w - f23.y) * kF16LoScale));
          // predicated stores instead of a branch around split + stores (the compiler sank the whole split into the branch:
          // BSSY / BRA / BSYNC per float4); a dead row's values never leave the registers and never reach the range check
          const uint32_t lv = (live >> it) & 1u;
          if (lv) mx2 = __hmax2(mx2, __hmax2(__habs2(g01), __habs2(g23)));
          st_cs_v2_if(hi16 + o, h01, h23, lv);
          st_cs_v2_if(lo16 + o, lo2.x, lo2.y, lv);
        } else if ((live >> it) & 1u) {
          store_split4_lob(ep.out_hi + split_base + o, lo16 + o, tv[it]);
        }
      }
    }
    __syncwarp();
  }
  if (LOB == 3) {
    // |hi| is inf exactly when the fp32 value was beyond fp16's range (round-to-nearest keeps everything below 65520 finite)
    if (__hisinf(__low2half(mx2)) || __hisinf(__high2half(mx2))) g_f16_overflow = 1;
  }
}

// ---- probe: does a SWIZZLE_128B K-major operand descriptor work when its start address is shifted by a whole
// number of 128-byte rows (not a multiple of 8)? One CTA, one 128 x 64 tile, plain single-pass TF32:
//   out[m][n] = sum_k A[m + shift][k] * W[n][k],   m in [0,128), K % 32 == 0, A has >= 128 + 8 rows.
// base_mode 0: descriptor base_offset field left 0; 1: base_offset = (start >> 7) & 7.
__global__ void __launch_bounds__(128, 1)
tc_shift_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW, int K, int shift,
                      int base_mode, float* __restrict__ out) {
  constexpr int BN = 64;
  constexpr int A_ROWS = kBM + 8;                  // 136 rows x 128 B = 17 KB (a multiple of 1024)
  constexpr int A_BYTES = A_ROWS * kBK * 4;
  constexpr int W_BYTES = BN * kBK * 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + A_BYTES + W_BYTES);
  uint64_t* done = bar + 1;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(bar, 1);
    tc::mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(64));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = tc::make_idesc(kBM, BN);
    const int nkb = K / kBK;
    for (int kb = 0; kb < nkb; ++kb) {
      if (kb > 0) tc::mbar_wait(done, (uint32_t)(kb - 1) & 1u);     // previous MMAs have read the single stage
      tc::mbar_expect_tx(bar, A_BYTES + W_BYTES);
      tc::tma_load_2d(smem, &tmA, bar, kb * kBK, 0);
      tc::tma_load_2d(smem + A_BYTES, &tmW, bar, kb * kBK, 0);
      tc::mbar_wait(bar, (uint32_t)kb & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a0 = tc::smem_u32(smem) + (uint32_t)shift * 128u;
      const uint32_t w0 = tc::smem_u32(smem + A_BYTES);
#pragma unroll
      for (int k = 0; k < kBK / kUmmaK; ++k) {
        uint64_t ad = tc::make_smem_desc(a0 + k * 32);
        if (base_mode == 1) ad |= (uint64_t)(((a0 + k * 32) >> 7) & 7u) << 49;
        tc::umma_tf32(tmem_base, ad, tc::make_smem_desc(w0 + k * 32), idesc, (uint32_t)((kb | k) != 0));
      }
      tc::umma_commit(done);
    }
    tc::mbar_wait(done, (uint32_t)(nkb - 1) & 1u);
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  {
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int row = warp * 32 + lane;
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      tc::tmem_ld16(tmem_base + lane_off + (uint32_t)c0, r);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) out[(long long)row * BN + c0 + i] = __uint_as_float(r[i]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64));
  }
}

}  // namespace tc2
}  // namespace mimi
