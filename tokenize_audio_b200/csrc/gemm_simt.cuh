// fp32 SIMT "row-gather" GEMM: every conv / linear of the Mimi encoder on channels-last activations.
//
// In channels-last layout [rows, C_in] the im2col row of a causal conv output j is the CONTIGUOUS run
// of k*C_in floats that starts at input row j*stride - (k - stride) (MimiConv1d, modeling_mimi.py:331-351:
// left pad k-stride, right pad up to a whole stride), so a conv is the GEMM
//     out[j, n] = sum_kk A[(j*stride - pad) * C_in + kk] * Wt[kk, n],   kk = tau*C_in + ci
// with rows < 0 or >= L read as zero ("constant" pad) or clamped (the "replicate" stride-2 downsample).
// Linears are the k=1, stride=1 case. The epilogue fuses bias, GELU(erf), LayerScale and the residual
// add; ELU on the conv INPUT (modeling_mimi.py:437-451) is applied while staging A.
//
// This is the exact-fp32 path (FFMA): the accuracy baseline of the tensor-core generations (mode 0) and the engine of the
// decode direction.
#pragma once
#include "common.cuh"

namespace mimi {

struct GemmParams {
  const float* A;             // [B][a_item_stride floats]; row r of item b at A + b*a_item_stride + r*Cin
  const float* Wt;            // [K][N], K index = tau*Cin + ci
  const float* bias;          // [N] or nullptr
  const float* scale;         // [N] LayerScale or nullptr
  const float* res;           // residual, same geometry as out, or nullptr (may alias out)
  float* out;                 // [B][out_item_stride floats]; row j at out + b*out_item_stride + j*N
  const int* len_in;          // device [B] valid input rows per item, or nullptr -> uniform_len_in
  int uniform_len_in;
  long long a_item_stride;
  long long out_item_stride;
  int Cin, stride, pad_left;  // conv geometry (linear: Cin=K, stride=1, pad_left=0)
  int K, N;
  int replicate;              // 1: clamp rows (replicate padding) instead of zero fill
  int elu_in;                 // 1: ELU applied to A while loading
  int act;                    // 0 none, 1 GELU(erf) after bias
};

template <int BM, int BN, int TN>
__global__ void __launch_bounds__(256) gemm_f32_kernel(const GemmParams p) {
  constexpr int BK = 16;
  constexpr int TM = 8;
  constexpr int NTX = BN / TN;
  constexpr int NTY = BM / TM;
  static_assert(NTX * NTY == 256, "256 threads per CTA");
  static_assert(TN == 4 || TN == 8, "TN");
  constexpr int LDA = BM + 4, LDB = BN + 4;
  constexpr int A_F4 = BM * BK / 4;             // float4 per A tile
  constexpr int B_F4 = BK * BN / 4;
  constexpr int A_LD = (A_F4 + 255) / 256;
  constexpr int B_LD = (B_F4 + 255) / 256;

  __shared__ __align__(16) float As[2][BK][LDA];
  __shared__ __align__(16) float Bs[2][BK][LDB];

  const int b = blockIdx.z;
  const int Lin = p.len_in ? p.len_in[b] : p.uniform_len_in;
  const int Lout = (Lin + p.stride - 1) / p.stride;
  const int j0 = blockIdx.x * BM;
  if (j0 >= Lout) return;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const float* __restrict__ Ab = p.A + (long long)b * p.a_item_stride;
  const float* __restrict__ Wt = p.Wt;

  float4 ra[A_LD], rb[B_LD];

  auto load_tile = [&](int k0) {
    const int tau = k0 / p.Cin;
    const int ci0 = k0 - tau * p.Cin;
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * 256;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < A_F4) {
        const int r = idx >> 2, kq = idx & 3;
        const int j = j0 + r;
        int row = j * p.stride - p.pad_left + tau;
        bool ok = j < Lout;
        if (p.replicate) row = min(max(row, 0), Lin - 1);
        else ok = ok && row >= 0 && row < Lin;
        if (ok) {
          v = ld_nc_f4(Ab + (long long)row * p.Cin + ci0 + kq * 4);
          if (p.elu_in) { v.x = elu1(v.x); v.y = elu1(v.y); v.z = elu1(v.z); v.w = elu1(v.w); }
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * 256;
      if (idx < B_F4) {
        const int kk = idx / (BN / 4), c4 = idx % (BN / 4);
        rb[i] = ld_nc_f4(Wt + (long long)(k0 + kk) * p.N + n0 + c4 * 4);
      }
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_LD; ++i) {
      const int idx = tid + i * 256;
      if (idx < A_F4) {
        const int r = idx >> 2, kq = idx & 3;
        As[buf][kq * 4 + 0][r] = ra[i].x;
        As[buf][kq * 4 + 1][r] = ra[i].y;
        As[buf][kq * 4 + 2][r] = ra[i].z;
        As[buf][kq * 4 + 3][r] = ra[i].w;
      }
    }
#pragma unroll
    for (int i = 0; i < B_LD; ++i) {
      const int idx = tid + i * 256;
      if (idx < B_F4) {
        const int kk = idx / (BN / 4), c4 = idx % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][kk][c4 * 4]) = rb[i];
      }
    }
  };

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int nk = p.K / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], bb[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][BM / 2 + ty * 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w;
      if constexpr (TN == 8) {
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][BN / 2 + tx * 4]);
        bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(cur ^ 1);
    __syncthreads();
  }

  // epilogue: bias -> activation -> LayerScale -> residual, float4 stores (coalesced along N)
  const long long obase = (long long)b * p.out_item_stride;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int r = (i < 4) ? (ty * 4 + i) : (BM / 2 + ty * 4 + (i - 4));
    const int j = j0 + r;
    if (j >= Lout) continue;
#pragma unroll
    for (int g = 0; g < TN / 4; ++g) {
      const int c = n0 + (g == 0 ? tx * 4 : BN / 2 + tx * 4);
      float4 v = make_float4(acc[i][g * 4 + 0], acc[i][g * 4 + 1], acc[i][g * 4 + 2], acc[i][g * 4 + 3]);
      if (p.bias) {
        const float4 bv = ld_nc_f4(p.bias + c);
        v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
      }
      if (p.act == 1) { v.x = gelu_erf(v.x); v.y = gelu_erf(v.y); v.z = gelu_erf(v.z); v.w = gelu_erf(v.w); }
      if (p.scale) {
        const float4 sv = ld_nc_f4(p.scale + c);
        v.x *= sv.x; v.y *= sv.y; v.z *= sv.z; v.w *= sv.w;
      }
      const long long o = obase + (long long)j * p.N + c;
      if (p.res) {
        const float4 rv = *reinterpret_cast<const float4*>(p.res + o);   // may alias out: plain load
        v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
      }
      *reinterpret_cast<float4*>(p.out + o) = v;
    }
  }
}

// L0: Conv1d(1 -> 64, k=7, stride 1) straight from the waveform (modeling_mimi.py:456, layers.0).
// x [B][x_stride] fp32, out channels-last [B][out_item_stride], row t = 64 floats. HBM-write bound.
__global__ void __launch_bounds__(256) conv0_kernel(const float* __restrict__ x, long long x_stride,
                                                    const float* __restrict__ w,     // [64][7]
                                                    const float* __restrict__ bias,  // [64]
                                                    float* __restrict__ out, long long out_item_stride,
                                                    const int* __restrict__ len_in, int uniform_len) {
  constexpr int TT = 128;                     // time steps per CTA
  __shared__ float xs[TT + 6];
  const int b = blockIdx.y;
  const int L = len_in ? len_in[b] : uniform_len;
  const int t0 = blockIdx.x * TT;
  if (t0 >= L) return;
  const int tid = threadIdx.x;
  const float* xb = x + (long long)b * x_stride;
  for (int i = tid; i < TT + 6; i += 256) {
    const int t = t0 - 6 + i;
    xs[i] = (t >= 0 && t < L) ? __ldg(xb + t) : 0.f;
  }
  const int cg = tid & 15;                    // channels cg*4 .. cg*4+3
  const int tl = tid >> 4;                    // 16 time lanes
  float wr[4][7], br[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    br[c] = __ldg(bias + cg * 4 + c);
#pragma unroll
    for (int k = 0; k < 7; ++k) wr[c][k] = __ldg(w + (cg * 4 + c) * 7 + k);
  }
  __syncthreads();
  float* ob = out + (long long)b * out_item_stride;
#pragma unroll
  for (int i = 0; i < TT / 16; ++i) {
    const int tloc = tl + 16 * i;
    const int t = t0 + tloc;
    if (t >= L) continue;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      const float xv = xs[tloc + k];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = fmaf(wr[c][k], xv, v[c]);
    }
    float4 o = make_float4(v[0] + br[0], v[1] + br[1], v[2] + br[2], v[3] + br[3]);
    *reinterpret_cast<float4*>(ob + (long long)t * 64 + cg * 4) = o;
  }
}

}  // namespace mimi
