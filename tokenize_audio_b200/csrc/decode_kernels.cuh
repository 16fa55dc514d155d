// Small kernels of the decode direction (codes -> waveform; MimiModel.decode, modeling_mimi.py:1613-1679). The convolutions,
// linears and the transformer run on the fp32 FFMA kernels of the encoder's exact-fp32 generation (gemm_simt.cuh,
// transformer.cuh): the decode direction is the reference's spot-check path (REF/*/utils.py:72-81), not a throughput path.
#pragma once
#include "common.cuh"

namespace mimi {

// MimiSplitResidualVectorQuantizer.decode up to the output projections (modeling_mimi.py:1282-1297, 1340-1350):
// q[b, t, 0:256] = embed_sem[codes[b,0,t]],  q[b, t, 256:512] = sum_{s>=1} embed_aco[s-1][codes[b,s,t]] (summed in stage order,
// as the reference's loop does). One warp per frame; codes outside [0, 2048) raise `bad` (the reference's F.embedding
// would fail on them).
__global__ void __launch_bounds__(256) rvq_decode_sum_kernel(const long long* __restrict__ codes, int B, int K, long long T,
                                                             const float* __restrict__ embed /*[32][2048][256]*/,
                                                             float* __restrict__ q /*[B*T][512]*/, int* __restrict__ bad) {
  const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (f >= (long long)B * T) return;
  const long long b = f / T, t = f - b * T;
  float4 sem[2], aco[2];
  for (int i = 0; i < 2; ++i) { sem[i] = make_float4(0.f, 0.f, 0.f, 0.f); aco[i] = sem[i]; }
  for (int s = 0; s < K; ++s) {
    const long long c = codes[(b * K + s) * T + t];
    if (c < 0 || c >= kCodebookSize) { if (lane == 0) *bad = 1; continue; }
    const float* e = embed + ((long long)s * kCodebookSize + c) * kCodeDim;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float4 v = ld_nc_f4(e + (i * 32 + lane) * 4);
      float4& d = s == 0 ? sem[i] : aco[i];
      d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w;
    }
  }
  float* o = q + f * 512;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    *reinterpret_cast<float4*>(o + (i * 32 + lane) * 4) = sem[i];
    *reinterpret_cast<float4*>(o + 256 + (i * 32 + lane) * 4) = aco[i];
  }
}

// MimiModel.upsample: depthwise ConvTranspose1d(512, 512, k = 4, stride 2, groups = 512, no bias), causal (the last k - s = 2
// outputs are trimmed; modeling_mimi.py:354-409, 1433-1441): y[2i + phi, c] = x[i, c] * w[c, phi] + x[i-1, c] * w[c, phi + 2].
__global__ void __launch_bounds__(256) upsample2_depthwise_kernel(const float* __restrict__ x /*[B][T][512]*/, long long T,
                                                                  const float* __restrict__ w /*[512][4]*/,
                                                                  float* __restrict__ y /*[B][2T][512]*/, long long n_out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;     // over B * 2T * 512
  if (i >= n_out) return;
  const int c = (int)(i % 512);
  const long long row = i / 512;                                            // b * 2T + (2 ti + phi)
  const long long b = row / (2 * T), r = row - b * 2 * T;
  const long long ti = r >> 1;
  const int phi = (int)(r & 1);
  const float* xb = x + b * T * 512;
  float acc = xb[ti * 512 + c] * __ldg(w + c * 4 + phi);
  if (ti > 0) acc = fmaf(xb[(ti - 1) * 512 + c], __ldg(w + c * 4 + phi + 2), acc);
  y[i] = acc;
}

// the last conv has one output channel, computed as column 0 of a 32-column GEMM: gather it into audio_values [B][1][L]
__global__ void __launch_bounds__(256) take_column0_kernel(const float* __restrict__ x /*[n][32]*/, float* __restrict__ y, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = x[i * 32];
}

}  // namespace mimi
