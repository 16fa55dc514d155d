// tcgen05 implicit-GEMM for the wide layers: 3xTF32 split precision, TMA-fed, accumulators in TMEM.
//
//   D[128 x BN] (fp32, TMEM) = A_hi*W_hi + A_hi*W_lo + A_lo*W_hi        (kind::tf32, UMMA 128 x BN x 8)
//
// A is the channels-last activation seen through a 3-D TMA tensor map whose row stride is
// conv_stride*C_in floats and whose inner extent is k*C_in floats (overlapping rows: the im2col matrix is
// never materialised; the causal left pad and the right "extra" pad are zero halo rows of the buffer).
// Both operands are stored pre-split: hi = fp32 rounded to the 10-bit TF32 mantissa, lo = x - hi (exact), so
// the three passes recover ~2^-22 relative accuracy per product whatever rounding the tensor core applies.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> bias / GELU / LayerScale / residual / ELU / hi-lo split -> global).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mimi {
namespace tc {

constexpr int kBM = 128;
constexpr int kBK = 32;                       // fp32 elements per k-block = one 128-byte swizzle row
constexpr int kUmmaK = 8;                     // tf32 MMA K

struct Epilogue {
  const float* cmul;            // [N] per-column factor: the weight unscaling of mode 9 (a power of two) x LayerScale, 1 otherwise
  const float* cadd;            // [N] per-column addend: bias x LayerScale, 0 otherwise:   out = act(acc * cmul + cadd) (+ res)
  const float* res;             // raw residual rows of N floats (may alias out_raw) or nullptr
  float* out_raw;               // raw output rows of N floats, or nullptr
  float* out_hi;                // split output (with halo rows), or nullptr
  float* out_lo;
  long long raw_item_stride;    // floats between items in res / out_raw
  long long split_item_stride;  // floats between items in out_hi / out_lo
  int split_front;              // halo rows in front of row 0 of each item in the split buffers
  int act;                      // 1: GELU(erf) after the affine through erff; 2: the same function through gelu_fast (common.cuh)
  int elu_split;                // 1: ELU applied before the hi/lo split (next conv's input activation)
  const int* len_in;            // device [B] input rows per item or nullptr -> uniform_len_in
  int uniform_len_in;
  int conv_stride;              // Lout = ceil(len_in / conv_stride)
  int N;
  int chunk_kb;                 // experiment: k-blocks per accumulation chunk (0 -> kChunkKB)
  int lo_bf16;                  // 1: out_lo is a bf16 array (mode 7), same element indexing as out_hi; 3: out_hi and out_lo are
                                //    fp16 arrays in the split_f16 format (mode 9)
  // Flattened linears (the B items as one [B * flat_rows][C] matrix, see tc_host.inl): rows past an item's length hold
  // whatever an earlier call left there. They are computed (rows are independent) but neither stored nor allowed to raise
  // the fp16 range flag: row r belongs to item r / flat_rows and is real iff r % flat_rows < flat_len[item].
  const int* flat_len;          // device [B] or nullptr (every row real)
  int flat_rows;                // 0: not a flattened launch
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// one lane of a converged warp (cute::elect_one_sync); lets the compiler keep single-thread tcgen05 issue code on
// the uniform datapath instead of wrapping every instruction in a divergence loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// warp index the compiler can prove warp-uniform (cutlass::canonical_warp_idx_sync)
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// pull a box towards L2 only (no shared-memory destination, no barrier)
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp):
// start address >> 4 in [0,14), LBO in [16,30) (unused for swizzled K-major), SBO >> 4 in [32,46) = 1024 B
// (8 rows x 128 B), version = 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10),
// both K-major, N >> 3 @ bit 17, M >> 4 @ bit 24.
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Cheap-to-issue form: the descriptor's high word is a constant (SBO = 1024 B, version 1, SWIZZLE_128B) and its low
// word is linear in the shared-memory address, so a k-step / operand / row shift is ONE integer add on a precomputed
// low word instead of re-encoding the address (the encode sequence on the uniform datapath made the single issuing
// thread the bottleneck: ~16 dependent uniform instructions per tcgen05.mma, profiles/r01_mma_issue.md).
constexpr uint32_t kDescHi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_tf32_lo(uint32_t tmem_d, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(kDescHi)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kChunkKB = 4;                   // k-blocks (of 32) accumulated inside TMEM before a drain: K = 128

// Accuracy note. The tensor core adds into its fp32 accumulator with truncation, so a long K loop drifts by
// ~0.3 ulp per MMA (measured: 3e-6 relative at K=512 growing to 1e-4 by K=8192, far above the 3e-5 the codes
// tolerate). Hence the GEMM kernels accumulate only kChunkKB k-blocks at a time inside TMEM: the epilogue warps drain each
// chunk (double-buffered) and add it to per-thread fp32 running sums with round-to-nearest FADDs while the next chunk is
// being computed.

// Zero the halo rows of one split buffer pair: rows [0, front) and [front + L_b, front + L_b + back) of every
// item (the causal left pad and the right "extra" pad of the consuming conv).
__global__ void zero_halo_kernel(float* hi, float* lo, long long item_stride, int C, int front, int back,
                                 const int* __restrict__ len, int uniform_len, int lob) {
  const int b = blockIdx.y;
  const int L = len ? len[b] : uniform_len;
  const int per = (front + back) * C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per; i += gridDim.x * blockDim.x) {
    int r = i / C;
    const int c = i - r * C;
    if (r >= front) r = front + L + (r - front);
    const long long o = (long long)b * item_stride + (long long)r * C + c;
    if (lob == 3) reinterpret_cast<uint16_t*>(hi)[o] = 0;
    else hi[o] = 0.f;
    reinterpret_cast<uint16_t*>(lo)[o] = 0;
  }
}

// z [B][rows][512] raw -> replicate-padded split copy for the stride-2 downsample conv
// (pad_mode="replicate", modeling_mimi.py:1422-1431): zp row 0,1 = z row 0; zp row 2+t = z row t;
// zp row 2+T = z row T-1.  One warp per zp row.
__global__ void __launch_bounds__(256) pad_replicate_split_kernel(const float* __restrict__ z, long long z_item_stride,
                                                                  float* __restrict__ hi, float* __restrict__ lo,
                                                                  long long split_item_stride,
                                                                  const int* __restrict__ len, int uniform_len, int lob) {
  const int b = blockIdx.y;
  const int T = len ? len[b] : uniform_len;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + warp;            // zp row
  if (T <= 0 || r >= T + 3) return;
  const int src = min(max(r - 2, 0), T - 1);
  const float* zr = z + (long long)b * z_item_stride + (long long)src * kHidden;
  const long long o = (long long)b * split_item_stride + (long long)r * kHidden;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    store_split4_x(hi, lo, o + c, ld_nc_f4(zr + c), lob);
  }
}

}  // namespace tc
}  // namespace mimi
