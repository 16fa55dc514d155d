// Fused split residual vector quantiser (MimiSplitResidualVectorQuantizer.encode, modeling_mimi.py:1311-1338
// over MimiResidualVectorQuantizer.encode :1262-1280 and MimiEuclideanCodebook.quantize :1197-1202).
//
// One CTA owns 64 frames and walks all K stages with the residual resident in shared memory: per stage a
// [64 x 2048 x 256] distance GEMM (codebook streamed k-major through smem), the torch.cdist formula
// sqrt(max(|x|^2 + |e|^2 - 2 x.e, 0)) with lowest-index tie-break, a cross-thread argmin, and the
// residual update r -= E[idx]. The 2048-wide distance rows never leave registers.
#pragma once
#include "common.cuh"

namespace mimi {

constexpr int kRvqFM = 64;                    // frames per CTA
constexpr int kRvqBN = 128;                   // codes per chunk
constexpr int kRvqBK = 16;
constexpr int kRvqLDR = kRvqFM + 4;           // residual pitch (k-major: R[k][frame])
constexpr int kRvqLDB = kRvqBN + 4;
constexpr size_t kRvqSmemBytes = sizeof(float) * (size_t)(kCodeDim * kRvqLDR + 2 * kRvqBK * kRvqLDB + kRvqFM * 4 +
                                                          kRvqFM + kRvqFM * 16) +
                                 sizeof(int) * (size_t)(kRvqFM * 16 + kRvqFM * 3);

struct RvqParams {
  const float* rproj;        // [B][item_stride]: row t = [P_sem e (256) | P_aco e (256)]
  long long item_stride;
  const float* embed;        // [32][2048][256] row-major (gather for the residual update)
  const float* embed_t;      // [32][256][2048] k-major (streamed for the distance GEMM)
  const float* enorm;        // [32][2048] |e|^2
  long long* codes;          // [B][K][T_out] int64
  int K, T_out;
  const int* len;            // device [B] frames per item or nullptr -> uniform_len
  int uniform_len;
  int B;
  int total_frames;          // sum of frames (uniform: B*uniform_len)
  const int* frame_prefix;   // device [B+1] prefix sums of len (ragged) or nullptr
};

__global__ void __launch_bounds__(256) rvq_encode_kernel(const RvqParams p) {
  extern __shared__ __align__(16) float smem[];
  float* R = smem;                                   // [256][68]
  float* Bs = R + kCodeDim * kRvqLDR;                // [2][16][132]
  float* part = Bs + 2 * kRvqBK * kRvqLDB;           // [4][64] partial |x|^2
  float* xn = part + kRvqFM * 4;                     // [64]
  float* cand_d = xn + kRvqFM;                       // [64][16]
  int* cand_i = reinterpret_cast<int*>(cand_d + kRvqFM * 16);   // [64][16]
  int* fr_b = cand_i + kRvqFM * 16;                  // [64] item of frame
  int* fr_t = fr_b + kRvqFM;                         // [64] frame index inside item
  int* best = fr_t + kRvqFM;                         // [64] chosen code

  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * kRvqFM;
  const int nf = min(kRvqFM, p.total_frames - f0);
  if (nf <= 0) return;

  if (tid < kRvqFM) {
    int b = 0, t = 0;
    const int m = f0 + tid;
    if (tid < nf) {
      if (p.frame_prefix) {
        int lo = 0, hi = p.B;                        // largest b with prefix[b] <= m
        while (hi - lo > 1) {
          const int mid = (lo + hi) >> 1;
          if (p.frame_prefix[mid] <= m) lo = mid; else hi = mid;
        }
        b = lo; t = m - p.frame_prefix[lo];
      } else {
        b = m / p.uniform_len; t = m - b * p.uniform_len;
      }
    }
    fr_b[tid] = b; fr_t[tid] = t;
  }
  __syncthreads();

  const int tx = tid & 15, ty = tid >> 4;            // 16 x 16 threads: 4 frames x 8 codes each

  for (int stage = 0; stage < p.K; ++stage) {
    // (re)load the residual: stage 0 = semantic projection, stage 1 = acoustic projection (the acoustic
    // chain restarts from the un-quantised latent, modeling_mimi.py:1330-1336)
    if (stage <= 1) {
      __syncthreads();
      for (int idx = tid; idx < kRvqFM * (kCodeDim / 4); idx += 256) {
        const int f = idx >> 6, k4 = (idx & 63) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (f < nf)
          v = ld_nc_f4(p.rproj + (long long)fr_b[f] * p.item_stride + (long long)fr_t[f] * 512 + stage * 256 + k4);
        R[(k4 + 0) * kRvqLDR + f] = v.x;
        R[(k4 + 1) * kRvqLDR + f] = v.y;
        R[(k4 + 2) * kRvqLDR + f] = v.z;
        R[(k4 + 3) * kRvqLDR + f] = v.w;
      }
    }
    __syncthreads();
    // |x|^2 per frame
    {
      const int f = tid & 63, q = tid >> 6;
      float s = 0.f;
      for (int k = q * 64; k < q * 64 + 64; ++k) { const float v = R[k * kRvqLDR + f]; s = fmaf(v, v, s); }
      part[q * kRvqFM + f] = s;
    }
    __syncthreads();
    if (tid < kRvqFM) xn[tid] = (part[tid] + part[kRvqFM + tid]) + (part[2 * kRvqFM + tid] + part[3 * kRvqFM + tid]);
    __syncthreads();

    const float* Et = p.embed_t + (long long)stage * kCodeDim * kCodebookSize;
    const float* en = p.enorm + (long long)stage * kCodebookSize;
    float bd[4];
    int bi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { bd[i] = INFINITY; bi[i] = 0; }
    float xnr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xnr[i] = xn[ty * 4 + i];

    for (int chunk = 0; chunk < kCodebookSize / kRvqBN; ++chunk) {
      const int c0 = chunk * kRvqBN;
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      float4 rb[2];
      auto load_b = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int idx = tid + i * 256;             // 16 rows x 32 float4
          const int kk = idx >> 5, c4 = idx & 31;
          rb[i] = ld_nc_f4(Et + (long long)(k0 + kk) * kCodebookSize + c0 + c4 * 4);
        }
      };
      auto store_b = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int idx = tid + i * 256;
          const int kk = idx >> 5, c4 = idx & 31;
          *reinterpret_cast<float4*>(Bs + (buf * kRvqBK + kk) * kRvqLDB + c4 * 4) = rb[i];
        }
      };
      load_b(0);
      store_b(0);
      __syncthreads();
      constexpr int NK = kCodeDim / kRvqBK;          // 16
      for (int kt = 0; kt < NK; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < NK) load_b((kt + 1) * kRvqBK);
#pragma unroll
        for (int kk = 0; kk < kRvqBK; ++kk) {
          const float4 a = *reinterpret_cast<const float4*>(R + (kt * kRvqBK + kk) * kRvqLDR + ty * 4);
          const float* brow = Bs + (cur * kRvqBK + kk) * kRvqLDB;
          const float4 b0 = *reinterpret_cast<const float4*>(brow + tx * 4);
          const float4 b1 = *reinterpret_cast<const float4*>(brow + kRvqBN / 2 + tx * 4);
          const float av[4] = {a.x, a.y, a.z, a.w};
          const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < NK) store_b(cur ^ 1);
        __syncthreads();
      }
      // distances for this chunk; codes visited in ascending order per thread -> strict '<' keeps the
      // lowest index among equal minima (torch argmin semantics)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int code = c0 + (j < 4 ? tx * 4 + j : kRvqBN / 2 + tx * 4 + (j - 4));
        const float e2 = __ldg(en + code);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float d2 = fmaf(-2.f, acc[i][j], xnr[i] + e2);
          const float d = sqrtf(fmaxf(d2, 0.f));
          if (d < bd[i]) { bd[i] = d; bi[i] = code; }
        }
      }
    }
    // cross-thread argmin over the 16 threads that share a frame
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      cand_d[(ty * 4 + i) * 16 + tx] = bd[i];
      cand_i[(ty * 4 + i) * 16 + tx] = bi[i];
    }
    __syncthreads();
    if (tid < kRvqFM) {
      float d = cand_d[tid * 16];
      int c = cand_i[tid * 16];
      for (int u = 1; u < 16; ++u) {
        const float du = cand_d[tid * 16 + u];
        const int cu = cand_i[tid * 16 + u];
        if (du < d || (du == d && cu < c)) { d = du; c = cu; }
      }
      best[tid] = c;
      if (tid < nf)
        p.codes[((long long)fr_b[tid] * p.K + stage) * p.T_out + fr_t[tid]] = (long long)c;
    }
    __syncthreads();
    // residual update r -= E[idx] (skipped after the semantic stage and after the last stage)
    if (stage >= 1 && stage + 1 < p.K) {
      const float* E = p.embed + (long long)stage * kCodebookSize * kCodeDim;
      for (int f = 0; f < kRvqFM; ++f) {
        const float e = __ldg(E + (long long)best[f] * kCodeDim + tid);   // tid = dim, coalesced row
        R[tid * kRvqLDR + f] -= e;
      }
    }
  }
}

// [B][T][512] channels-last latent -> [B][512][T] (MimiModel layout) for parity dumps only.
__global__ void latent_transpose_kernel(const float* __restrict__ e, long long item_stride,
                                        float* __restrict__ out, int T_out, const int* __restrict__ len,
                                        int uniform_len) {
  const int b = blockIdx.y;
  const int T = len ? len[b] : uniform_len;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over 512*T_out
  if (i >= (long long)kHidden * T_out) return;
  const int c = (int)(i / T_out), t = (int)(i % T_out);
  out[(long long)b * kHidden * T_out + i] = t < T ? e[(long long)b * item_stride + (long long)t * kHidden + c] : 0.f;
}

__global__ void fill_codes_zero_kernel(long long* codes, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) codes[i] = 0;
}

}  // namespace mimi
