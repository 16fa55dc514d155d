// Shared helpers for the mimi_b200 kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

namespace mimi {

constexpr int kHidden = 512;
constexpr int kHeads = 8;
constexpr int kHeadDim = 64;
constexpr int kWindow = 250;      // MimiConfig.sliding_window (modeling_mimi.py:1096-1102)
constexpr int kFfn = 2048;
constexpr int kCodebookSize = 2048;
constexpr int kCodeDim = 256;

// nn.ELU(alpha=1): x > 0 ? x : exp(x) - 1   (modeling_mimi.py:428,473,478)
__device__ __forceinline__ float elu1(float x) { return x > 0.f ? x : expm1f(x); }

// ELU for the tensor-core epilogues: exp(x) - 1 through ex2.approx (absolute error ~1e-7, below the fp32 rounding
// of the activations it feeds; end-to-end latent error is unchanged, see tests/test_gpu_parity.py)
__device__ __forceinline__ float elu_fast(float x) {
  float e, r = x;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
  // r = x > 0 ? x : e - 1 as one compare and one predicated add (left to itself the compiler sometimes emits add + compare +
  // select: 5 instead of 4 instructions per ELU, and the front end runs 160 of them per row)
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, 0f00000000;\n\t@!p add.f32 %0, %2, 0fBF800000;\n\t}" : "+f"(r) : "f"(x), "f"(e));
  return r;
}

// exact (erf) GELU, ACT2FN["gelu"] used by MimiMLP (modeling_mimi.py:614-627)
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
}

// The same GELU for the tensor-core epilogues, 15 instructions instead of erff's 25 (fc1 is bound by its tile finish):
// gelu(x) = x * Phi(x), Phi(-t) = 0.5 erfc(t / sqrt 2) = 2^q(t) with q a degree-10 polynomial fitted to log2(0.5 erfc(t / sqrt 2)) on
// [0, 5.5] (max error 2e-7 in q), Phi(t) = 1 - Phi(-t). Beyond 5.5 Phi is 1.9e-8 / 1 - 1.9e-8 either way. Against the exact
// function in double: absolute error <= 3.9e-7 over [-8, 8] (an erff-based fp32 GELU: 4.5e-7, both set by the rounding of the
// result), relative error <= 1.2e-6 for x > -3 (<= 8e-7 on [-2, 2]).
__device__ __forceinline__ float gelu_fast(float x) {
  const float t = fminf(fabsf(x), 5.5f);
  float q = -1.725302567479048e-08f;
  q = fmaf(q, t, 5.477549507304502e-07f);
  q = fmaf(q, t, -7.4582894740160555e-06f);
  q = fmaf(q, t, 5.484473149408586e-05f);
  q = fmaf(q, t, -0.000200506707187742f);
  q = fmaf(q, t, -0.00017614704847801477f);
  q = fmaf(q, t, 0.007180649787187576f);
  q = fmaf(q, t, -0.05261564627289772f);
  q = fmaf(q, t, -0.4591561555862427f);
  q = fmaf(q, t, -1.1511132717132568f);
  q = fmaf(q, t, -0.9999998211860657f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));
  return x * (x >= 0.f ? 1.f - e : e);
}

__device__ __forceinline__ float4 ld_nc_f4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// fp32 -> (hi, lo) for the 3xTF32 tensor-core path: hi keeps the 10-bit TF32 mantissa (round to nearest on
// the dropped 13 bits), lo = x - hi exactly. hi is exactly TF32-representable, so whatever rounding the
// tensor core applies to its fp32 inputs cannot change it.
__host__ __device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
#ifdef __CUDA_ARCH__
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
#else
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&hi, &u, 4);
#endif
  lo = x - hi;
}
// lo part of the split as bf16 (mode 7): lo = x - hi is a correction of relative size 2^-11, so 8 mantissa bits keep
// the pair (hi, lo) accurate to 2^-20 of x; the A_lo * W_hi product then runs on kind::f16 at twice the TF32 rate
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {      // a -> low half, b -> high half (round to nearest)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// hi: 4 floats at `hi`; lo: 4 bf16 at element offset of the same index in a bf16 array
__device__ __forceinline__ void store_split4_lob(float* hi, void* lo_bf16_elem, float4 v) {
  float4 h4, l4;
  split_tf32(v.x, h4.x, l4.x); split_tf32(v.y, h4.y, l4.y); split_tf32(v.z, h4.z, l4.z); split_tf32(v.w, h4.w, l4.w);
  *reinterpret_cast<float4*>(hi) = h4;
  *reinterpret_cast<uint2*>(lo_bf16_elem) = make_uint2(pack_bf16x2(l4.x, l4.y), pack_bf16x2(l4.z, l4.w));
}
// fp16 generation (mode 9, lob = 3): x = hi + lo / 2048 with hi = fp16(x) (11 significant bits) and lo = fp16((x - hi) * 2048)
// (the next 11 bits; the scaling keeps the correction in fp16's normal range however small x is), 4 bytes per element.
// Every product of the consuming GEMM -- hi*W_hi, hi*W_lo, lo*(W_hi / 2048) -- then runs on kind::f16 at the full 16-bit
// tensor rate: 3 passes instead of the 5 pass units of the TF32 scheme. fp16 ends at 65504: a value beyond it raises the range
// flag below (Mimi activations are O(1..1e2); the heavy-tailed parity fixture reaches 1e4).
constexpr float kF16LoScale = 2048.f;
// set (never cleared by a kernel) when a value outside fp16's range was met: the host reads it behind every mode-9
// encode (mimi_b200_range_overflow) and the wrapper re-encodes such a batch with the range-safe TF32 generation
__device__ int g_f16_overflow = 0;
__device__ __forceinline__ void split_f16(float x, float& hi, float& lo) {
  if (!(fabsf(x) <= 65504.f)) g_f16_overflow = 1;     // (also NaN); the value then becomes inf and the batch is re-encoded
  hi = __half2float(__float2half_rn(x));
  lo = (x - hi) * kF16LoScale;
}
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {       // a -> low half, b -> high half (round to nearest)
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// hi16 / lo16: element idx of two fp16 arrays that start at `hi` / `lo`. One range check per four values; hi comes out of
// the packed conversion (cvt.rn.f16x2 + two unpacks) instead of four scalar round trips.
__device__ __forceinline__ void store_split4_f16(float* hi, float* lo, long long idx, float4 v) {
  if (!(fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))) <= 65504.f)) g_f16_overflow = 1;
  const uint32_t h01 = pack_f16x2(v.x, v.y), h23 = pack_f16x2(v.z, v.w);
  const float2 f01 = __half22float2(*reinterpret_cast<const __half2*>(&h01));
  const float2 f23 = __half22float2(*reinterpret_cast<const __half2*>(&h23));
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi) + idx) = make_uint2(h01, h23);
  *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(lo) + idx) =
      make_uint2(pack_f16x2((v.x - f01.x) * kF16LoScale, (v.y - f01.y) * kF16LoScale),
                 pack_f16x2((v.z - f23.x) * kF16LoScale, (v.w - f23.y) * kF16LoScale));
}

// one entry point for every producer of a split activation: lob = 1 -> TF32 hi (fp32) + bf16 lo (mode 7), lob = 3 -> fp16 pair
__device__ __forceinline__ void store_split4_x(float* hi, float* lo, long long idx, float4 v, int lob) {
  if (lob == 3) store_split4_f16(hi, lo, idx, v);
  else store_split4_lob(hi + idx, reinterpret_cast<uint16_t*>(lo) + idx, v);
}

// explicit shared-space 16-byte accesses (a pointer into dynamic shared memory that went through integer
// arithmetic is otherwise compiled as a generic LD/ST, which the compiler cannot reorder against global
// stores and which pays the generic-address check)
__device__ __forceinline__ void sts128(uint32_t saddr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace mimi
