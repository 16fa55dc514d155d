// tcgen05 implicit-GEMM, second generation: PERSISTENT, with the tile epilogue overlapped with the next
// tile's main loop and coalesced global I/O. Same arithmetic as tc_gemm.cuh (3xTF32, chunked fp32
// accumulation, separate accumulator for the two small cross terms), same operands (pre-split hi/lo
// activations seen through overlapping-row TMA maps, pre-split K-major weights), same Epilogue contract.
//
//   grid   = min(#virtual tiles, #SMs) CTAs of 320 threads, one per SM, static round-robin over
//            virtual tiles id -> (m-tile, item, n-tile); tiles past an item's length are skipped by all roles.
//   warp 0 = TMA producer (runs ahead across tile boundaries through a 3..5-stage ring)
//   warp 1 = TMEM owner + single-thread tcgen05.mma issuer
//   warps 2..17 = epilogue (2..9 for BN = 32): four warps per TMEM lane quarter, each owning a quarter of the BN
//            columns -- the epilogue (chunk drains + ELU / split / stores) is what bounds most layers, and it is
//            instruction-latency bound, so it gets four warps per SM sub-partition. They drain
//            every K=128 chunk of hi*hi from TMEM (double-buffered) into fp32 registers (round-to-nearest
//            adds), then the cross-term accumulator (double-buffered per tile parity), then finish the tile:
//            bias / GELU / LayerScale per thread = per row, a 32-column transpose through a swizzled smem
//            staging tile, and row-contiguous float4 residual loads and raw / hi / lo stores (128 B segments).
//   TMEM   = 2 chunk buffers x [main | cross terms], 4*BN columns (BN <= 128).
#pragma once
#include "tc_gemm.cuh"

namespace mimi {
namespace tc2 {

using tc::Epilogue;
using tc::kBK;
using tc::kBM;
using tc::kChunkKB;
using tc::kUmmaK;

__host__ __device__ constexpr int epi_warps(int bn) { return bn >= 64 ? 16 : 8; }
__host__ __device__ constexpr int threads(int bn) { return 64 + 32 * epi_warps(bn); }       // 576 / 320
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

// i-th k-block to visit -> linear k-block index (tau * cp + channel panel)
__device__ __forceinline__ int kblock_order(const Sched& sc, int i) {
  if (sc.G <= 1) return i;
  if (sc.G >= 16) {
    // variant (G + 16): channel panels innermost -- the cp 128-byte pieces of one input row are fetched back to back
    // (DRAM page locality), the tap that re-reads the same rows follows cp k-blocks later (still an L2 hit)
    const int G = sc.G - 16;
    const int cb = i % sc.cp, t = i / sc.cp;
    const int dq = t % G, ph = t / G;
    return (ph + sc.s * dq) * sc.cp + cb;
  }
  const int dq = i % sc.G, t = i / sc.G;
  const int cb = t % sc.cp, ph = t / sc.cp;
  return (ph + sc.s * dq) * sc.cp + cb;
}

template <int BN>
struct Cfg {
  static constexpr int A_BYTES = kBM * kBK * 4;                      // 16 KB
  static constexpr int W_BYTES = BN * kBK * 4;
  static constexpr int STAGE = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr int EPIW = epi_warps(BN);
  static constexpr int PC = 16;                                      // staging piece width (columns)
  static constexpr int STG_WARP = 32 * PC * 4;                       // staging bytes per epilogue warp
  static constexpr int STG = EPIW * STG_WARP;
  static constexpr int BAR_BYTES = 512;
  static constexpr int STAGES_RAW = (kSmemMax - 1024 - STG - BAR_BYTES) / STAGE;
  static constexpr int STAGES = STAGES_RAW > 6 ? 6 : STAGES_RAW;
  static constexpr int SMEM = 1024 + STAGES * STAGE + STG + BAR_BYTES;
  static constexpr int TMEM_COLS = (4 * BN <= 32) ? 32 : (4 * BN <= 64) ? 64 : (4 * BN <= 128) ? 128 : (4 * BN <= 256) ? 256 : 512;
  static_assert(BN == 32 || BN == 64 || BN == 128, "BN");
  static_assert(STAGES >= 3, "ring too shallow");
};

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

// Finish a tile: thread (lane) holds acc[0, HALF) = columns ncol0 + [0, HALF) of row row_base + lane. Bias / GELU /
// LayerScale per thread = per row, transpose through the warp's staging tile `stg`, then row-contiguous float4 residual
// loads and raw / hi / lo stores (128-byte segments).
template <int HALF, int PC>
__device__ __forceinline__ void finish_tile(const Epilogue& ep, float (&acc)[HALF], int b, int row_base, int ncol0, int Lout,
                                            uint32_t stg, int lane) {
  using G = EpiGeom<HALF, PC>;
  constexpr int NP = G::NP, LPR = G::LPR, RPI = G::RPI, IT = G::IT;
  const int rr = lane / LPR;                     // coalesced phase: row within an RPI-row group
  const int cj = lane % LPR;                     //                  float4 index inside the PC-wide piece
  const long long raw_base = (long long)b * ep.raw_item_stride + ncol0;
  const long long split_base = (long long)b * ep.split_item_stride + (long long)ep.split_front * ep.N + ncol0;
  // rows of this thread's coalesced phase that exist (bit `it`): inside the tile's item, and -- in a flattened launch -- inside
  // the length of the item the row falls into
  uint32_t live = 0;
#pragma unroll
  for (int it = 0; it < IT; ++it) {
    const int row = row_base + it * RPI + rr;
    bool ok = row < Lout;
    if (ok && ep.flat_rows > 0 && ep.flat_len) ok = (row % ep.flat_rows) < __ldg(ep.flat_len + row / ep.flat_rows);
    live |= (uint32_t)ok << it;
  }
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float4 resv[IT];
    if (ep.res) {
#pragma unroll
      for (int it = 0; it < IT; ++it) {
        const int row = row_base + it * RPI + rr;
        resv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((live >> it) & 1u) resv[it] = *reinterpret_cast<const float4*>(ep.res + raw_base + (long long)row * ep.N + p * PC + cj * 4);
      }
    }
    const int wkey = (LPR == 8) ? (lane & 7) : ((lane >> 1) & (LPR - 1));
#pragma unroll
    for (int j = 0; j < LPR; ++j) {
      float4 v = make_float4(acc[p * PC + 4 * j], acc[p * PC + 4 * j + 1], acc[p * PC + 4 * j + 2], acc[p * PC + 4 * j + 3]);
      const int c = ncol0 + p * PC + 4 * j;
      if (ep.wscale) {
        const float4 t = ld_nc_f4(ep.wscale + c);
        v.x *= t.x; v.y *= t.y; v.z *= t.z; v.w *= t.w;
      }
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
      const int r = it * RPI + rr;               // row within this warp's 32
      const int rkey = (LPR == 8) ? (r & 7) : ((r >> 1) & (LPR - 1));
      tv[it] = lds128(stg + (uint32_t)(r * LPR + (cj ^ rkey)) * 16u);
    }
#pragma unroll
    for (int it = 0; it < IT; ++it) {
      const int r = it * RPI + rr;
      float4 v = tv[it];
      const int row = row_base + r;
      if ((live >> it) & 1u) {
        const long long o = (long long)row * ep.N + p * PC + cj * 4;
        if (ep.res) { v.x += resv[it].x; v.y += resv[it].y; v.z += resv[it].z; v.w += resv[it].w; }
        if (ep.out_raw) *reinterpret_cast<float4*>(ep.out_raw + raw_base + o) = v;
        if (ep.out_hi) {
          if (ep.elu_split) { v.x = elu_fast(v.x); v.y = elu_fast(v.y); v.z = elu_fast(v.z); v.w = elu_fast(v.w); }
          store_split4_x(ep.out_hi, ep.out_lo, ep.out_hib, split_base + o, v, ep.lo_bf16);
        }
      }
    }
    __syncwarp();
  }
}

// The epilogue role shared by the k-block-ring kernel (tc2_gemm_kernel) and the plane-staged kernel
// (tc2p_gemm_kernel): drain every accumulation chunk, add the cross-term accumulator, finish the tile.
template <int BN, int PC_, int EPIW, int EW0 = 2>
__device__ __forceinline__ void epilogue_role(const Epilogue& ep, const Sched& sc, uint8_t* stg_base, uint64_t* acc_full,
                                              uint64_t* acc_empty, uint32_t tmem_base, int nchunks,
                                              int warp, int lane) {
  constexpr int HALF = BN / (EPIW / 4);            // columns per epilogue thread
  constexpr int PC = PC_;
  constexpr int STG_WARP = 32 * PC * 4;
  const int vtiles = sched_tiles(sc) * sc.ntn;
  auto decode = [&](int id, int& b, int& m0, int& n0, int& Lout) {
    n0 = (id % sc.ntn) * BN;
    return sched_tile(sc, ep, id / sc.ntn, b, m0, Lout);
  };
  // ---- epilogue warps --------------------------------------------------------------------------------------
  const int ew = warp - EW0;                     // EW0 = index of the first epilogue warp (a multiple of 4 or 2)
  const int quarter = warp & 3;                  // TMEM lanes [32*quarter, +32) are the ones this warp may read
  const int half = ew >> 2;                      // which slice of the BN columns
  const int col0 = half * HALF;
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  const uint32_t stg = tc::smem_u32(stg_base + ew * STG_WARP);         // [32 rows][LPR float4], swizzled
  uint32_t cc = 0;
  for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
    int b, m0, n0, Lout;
    if (!decode(id, b, m0, n0, Lout)) continue;
    float acc[HALF];
#pragma unroll
    for (int i = 0; i < HALF; ++i) acc[i] = 0.f;
    prefetch_residual<HALF, PC>(ep, b, m0 + quarter * 32, n0 + col0, Lout, lane);
    for (int c = 0; c < nchunks; ++c, ++cc) {
      const uint32_t buf = cc & 1u;
      tc::mbar_wait(&acc_full[buf], (cc >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // buffer layout per chunk: [main (BN columns) | cross terms (BN columns)], see the MMA issuer
      drain_add<HALF>(tmem_base + lane_off + buf * (2 * BN) + (uint32_t)col0, acc);
      drain_add<HALF>(tmem_base + lane_off + buf * (2 * BN) + BN + (uint32_t)col0, acc);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
    }
    // ---- finish the tile: rows m0 + 32*quarter + [0,32), columns n0 + col0 + [0,HALF) -----------------------
    finish_tile<HALF, PC>(ep, acc, b, m0 + quarter * 32, n0 + col0, Lout, stg, lane);
  }
}

template <int BN>
__global__ void __launch_bounds__(threads(BN), 1)
tc2_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, int K,
                const Epilogue ep, const Sched sc) {
  using C = Cfg<BN>;
  constexpr int STAGES = C::STAGES;
  constexpr int STAGE = C::STAGE;
  constexpr int A_BYTES = C::A_BYTES;
  constexpr int W_BYTES = C::W_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + STAGES * STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg_base + C::STG);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_full = empty_bar + STAGES;         // [2] main chunk ready        (MMA -> epilogue)
  uint64_t* acc_empty = acc_full + 2;              // [2] main chunk drained      (epilogue -> MMA)
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int nkb = K / kBK;
  const int ckb = ep.chunk_kb > 0 ? ep.chunk_kb : kChunkKB;
  const int nchunks = (nkb + ckb - 1) / ckb;
  const int vtiles = sched_tiles(sc) * sc.ntn;

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi); tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmW_hi); tc::prefetch_tmap(&tmW_lo);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&full_bar[s], 1); tc::mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], C::EPIW);
    }
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

  // virtual tile id -> (n-tile, item, m-tile); returns false for tiles past the item's length
  auto decode = [&](int id, int& b, int& m0, int& n0, int& Lout) {
    n0 = (id % sc.ntn) * BN;
    return sched_tile(sc, ep, id / sc.ntn, b, m0, Lout);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t kbc = 0;
      for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
        int b, m0, n0, Lout;
        if (!decode(id, b, m0, n0, Lout)) continue;
        // the activation rows of this CTA's NEXT tile are pulled towards L2 one tile ahead, k-block by k-block, so
        // that their TMA loads hit L2 instead of paying the DRAM latency with only three stages in flight
        int pb = 0, pm0 = 0, pn0 = 0, pL = 0;
        bool pf = false;
        if (ep.prefetch_next)
          for (int nid = id + gridDim.x; nid < vtiles && !pf; nid += gridDim.x) pf = decode(nid, pb, pm0, pn0, pL);
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          if (pf) {
            tc::tma_prefetch_3d(&tmA_hi, kb * kBK, pm0, pb);
            tc::tma_prefetch_3d(&tmA_lo, kb * kBK, pm0, pb);
          }
          const uint32_t s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1u;
          tc::mbar_wait(&empty_bar[s], ph ^ 1u);
          uint8_t* st = smem + s * STAGE;
          tc::mbar_expect_tx(&full_bar[s], STAGE);
          const int kx = kblock_order(sc, kb) * kBK;
          tc::tma_load_3d(st, &tmA_hi, &full_bar[s], kx, m0, b);
          tc::tma_load_3d(st + A_BYTES, &tmA_lo, &full_bar[s], kx, m0, b);
          tc::tma_load_2d(st + 2 * A_BYTES, &tmW_hi, &full_bar[s], kx, n0);
          tc::tma_load_2d(st + 2 * A_BYTES + W_BYTES, &tmW_lo, &full_bar[s], kx, n0);
        }
      }
    }
  } else if (warp == 1) {
    {
      // The whole warp walks the loop and waits on the barriers (converged), one elected lane issues: this keeps
      // descriptor arithmetic on the uniform datapath (no per-instruction divergence loop around tcgen05.mma).
      // Two MMAs per k-step instead of three: the stage keeps W_hi and W_lo adjacent, so ONE N = 2*BN MMA with A_hi
      // yields A_hi W_hi^T (columns [0,BN): main) and A_hi W_lo^T (columns [BN,2BN): cross), and an N = BN MMA adds
      // A_lo W_hi^T to the cross columns. A_hi is read from shared memory once instead of twice -- operand reads
      // (128 B/clk/SM for three 128x128x8 MMAs) plus the TMA fills are what saturates shared memory here.
      constexpr uint32_t idesc = tc::make_idesc(kBM, BN);
      constexpr uint32_t idesc2 = tc::make_idesc(kBM, 2 * BN);
      const uint32_t smem_base_u32 = tc::smem_u32(smem);
      uint32_t kbc = 0, cc = 0;
      for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
        int b, m0, n0, Lout;
        if (!decode(id, b, m0, n0, Lout)) continue;
        for (int c = 0; c < nchunks; ++c, ++cc) {
          const uint32_t buf = cc & 1u;
          tc::mbar_wait(&acc_empty[buf], ((cc >> 1) & 1u) ^ 1u);         // drained two chunks ago
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t tmem_main = tmem_base + buf * (2 * BN);
          const int kb_end = min(nkb, (c + 1) * ckb);
          for (int kb = c * ckb; kb < kb_end; ++kb, ++kbc) {
            const uint32_t s = kbc % STAGES;
            const uint32_t ph = (kbc / STAGES) & 1u;
            tc::mbar_wait(&full_bar[s], ph);
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
  } else {
    epilogue_role<BN, C::PC, C::EPIW>(ep, sc, stg_base, acc_full, acc_empty, tmem_base, nchunks, warp, lane);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
  }
}

// ---- plane-staged variant for convs with kernel k = G * stride, G >= 2 -----------------------------------------------
// Same tile (128 x BN), same accumulators and the same epilogue as tc2_gemm_kernel, but the activation is not fetched
// once per tap: plane (ph, p) = input rows {q*stride + ph} x channels [32p, 32p+32) is staged ONCE per tile through a
// 4-D TMA map (box 32 x 1 x (128+G-1) x 1) and serves the G taps tau = ph + stride*dq through UMMA descriptors whose
// start is shifted by dq rows (profiles/r01_sw128_row_shift_probe.md). The activation share of the L2 -> shared-memory
// stream, which is what bounds these kernels, drops by G; the weights keep their own k-block ring.
struct PlaneGeom {
  int G;          // taps per plane = k / stride (2 or 3)
  int s;          // conv stride
  int cpanels;    // C_in / 32
};

template <int BN>
struct CfgP {
  static constexpr int A_HALF = 136 * 128;                           // 128 + G - 1 <= 130 rows, rounded to whole KB
  static constexpr int A_STAGE = 2 * A_HALF;                         // hi | lo
  static constexpr int NA = (BN == 128) ? 2 : 3;
  static constexpr int W_BYTES = BN * kBK * 4;
  static constexpr int W_STAGE = 2 * W_BYTES;
  static constexpr int EPIW = epi_warps(BN);
  static constexpr int PC = 16;
  static constexpr int STG = EPIW * 32 * PC * 4;
  static constexpr int BAR_BYTES = 512;
  static constexpr int NW_RAW = (kSmemMax - 1024 - NA * A_STAGE - STG - BAR_BYTES) / W_STAGE;
  static constexpr int NW = NW_RAW > 6 ? 6 : NW_RAW;
  static constexpr int SMEM = 1024 + NA * A_STAGE + NW * W_STAGE + STG + BAR_BYTES;
  static constexpr int TMEM_COLS = (4 * BN <= 256) ? 256 : 512;
  static_assert(BN == 64 || BN == 128, "BN");
  static_assert(NW >= 3, "weight ring too shallow");
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(tc::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(threads(BN), 1)
tc2p_gemm_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                 const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo,
                 const Epilogue ep, const Sched sc, const PlaneGeom gm) {
  using C = CfgP<BN>;
  constexpr int NA = C::NA, NW = C::NW;
  constexpr int A_STAGE = C::A_STAGE, A_HALF = C::A_HALF, W_STAGE = C::W_STAGE, W_BYTES = C::W_BYTES;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_base = smem + NA * A_STAGE;
  uint8_t* stg_base = w_base + NW * W_STAGE;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(stg_base + C::STG);
  uint64_t* a_empty = a_full + NA;
  uint64_t* w_full = a_empty + NA;
  uint64_t* w_empty = w_full + NW;
  uint64_t* acc_full = w_empty + NW;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  const int nplanes = gm.s * gm.cpanels;
  const int nkb = nplanes * gm.G;
  const int ckb = ep.chunk_kb > 0 ? ep.chunk_kb : kChunkKB;
  const int nchunks = (nkb + ckb - 1) / ckb;
  const int vtiles = sched_tiles(sc) * sc.ntn;
  const uint32_t a_bytes = (uint32_t)(2 * (128 + gm.G - 1) * 128);

  if (warp == 0 && lane == 0) {
    tc::prefetch_tmap(&tmA_hi); tc::prefetch_tmap(&tmA_lo); tc::prefetch_tmap(&tmW_hi); tc::prefetch_tmap(&tmW_lo);
    for (int s = 0; s < NA; ++s) { tc::mbar_init(&a_full[s], 1); tc::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < NW; ++s) { tc::mbar_init(&w_full[s], 1); tc::mbar_init(&w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&acc_full[s], 1);
      tc::mbar_init(&acc_empty[s], C::EPIW);
    }
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
    return sched_tile(sc, ep, id / sc.ntn, b, m0, Lout);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t ac = 0, wc = 0;
      for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
        int b, m0, n0, Lout;
        if (!decode(id, b, m0, n0, Lout)) continue;
        for (int pi = 0; pi < nplanes; ++pi, ++ac) {
          const int ph = pi / gm.cpanels, pn = pi - ph * gm.cpanels;
          const uint32_t as = ac % NA;
          tc::mbar_wait(&a_empty[as], ((ac / NA) & 1u) ^ 1u);
          uint8_t* st = smem + as * A_STAGE;
          tc::mbar_expect_tx(&a_full[as], a_bytes);
          tma_load_4d(st, &tmA_hi, &a_full[as], pn * 32, ph, m0, b);
          tma_load_4d(st + A_HALF, &tmA_lo, &a_full[as], pn * 32, ph, m0, b);
          for (int dq = 0; dq < gm.G; ++dq, ++wc) {
            const int kbw = (ph + gm.s * dq) * gm.cpanels + pn;     // weight k-block of tap ph + s*dq, panel pn
            const uint32_t ws = wc % NW;
            tc::mbar_wait(&w_empty[ws], ((wc / NW) & 1u) ^ 1u);
            uint8_t* wt = w_base + ws * W_STAGE;
            tc::mbar_expect_tx(&w_full[ws], W_STAGE);
            tc::tma_load_2d(wt, &tmW_hi, &w_full[ws], kbw * kBK, n0);
            tc::tma_load_2d(wt + W_BYTES, &tmW_lo, &w_full[ws], kbw * kBK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    {
      constexpr uint32_t idesc = tc::make_idesc(kBM, BN);
      constexpr uint32_t idesc2 = tc::make_idesc(kBM, 2 * BN);
      const uint32_t smem_base_u32 = tc::smem_u32(smem), w_base_u32 = tc::smem_u32(w_base);
      uint32_t ac = 0, wc = 0, cc = 0;
      for (int id = blockIdx.x; id < vtiles; id += gridDim.x) {
        int b, m0, n0, Lout;
        if (!decode(id, b, m0, n0, Lout)) continue;
        int kb = 0;
        for (int pi = 0; pi < nplanes; ++pi, ++ac) {
          const uint32_t as = ac % NA;
          tc::mbar_wait(&a_full[as], (ac / NA) & 1u);
          const uint32_t d_a0 = tc::desc_lo(smem_base_u32 + as * A_STAGE);
          for (int dq = 0; dq < gm.G; ++dq, ++wc, ++kb) {
            const uint32_t buf = cc & 1u;
            const bool first_in_chunk = (kb % ckb) == 0;
            if (first_in_chunk) tc::mbar_wait(&acc_empty[buf], ((cc >> 1) & 1u) ^ 1u);
            const uint32_t ws = wc % NW;
            tc::mbar_wait(&w_full[ws], (wc / NW) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tmem_main = tmem_base + buf * (2 * BN);
            const uint32_t d_ahi = d_a0 + (uint32_t)(dq * 8);                 // dq rows of 128 B, in 16-byte units
            const uint32_t d_alo = d_ahi + (A_HALF >> 4);
            const uint32_t d_whi = tc::desc_lo(w_base_u32 + ws * W_STAGE);    // [W_hi | W_lo] adjacent: N = 2*BN
            const bool chunk_end = (kb + 1) % ckb == 0 || kb + 1 == nkb;
            if (tc::elect_one()) {
#pragma unroll
              for (int k = 0; k < kBK / kUmmaK; ++k) {
                tc::umma_tf32_lo(tmem_main, d_ahi + 2 * k, d_whi + 2 * k, idesc2, !(first_in_chunk && k == 0));
                tc::umma_tf32_lo(tmem_main + BN, d_alo + 2 * k, d_whi + 2 * k, idesc, 1u);
              }
              tc::umma_commit(&w_empty[ws]);
              if (chunk_end) tc::umma_commit(&acc_full[buf]);
              if (dq + 1 == gm.G) tc::umma_commit(&a_empty[as]);
            }
            __syncwarp();
            if (chunk_end) ++cc;
          }
        }
      }
    }
  } else {
    epilogue_role<BN, C::PC, C::EPIW>(ep, sc, stg_base, acc_full, acc_empty, tmem_base, nchunks, warp, lane);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(C::TMEM_COLS));
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
