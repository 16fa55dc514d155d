// The two HBM-bound byte/sample kernels either side of the encoder:
//  * polyphase FIR resampler to 24 kHz   (stands in for librosa.resample, REF/*/utils.py:84-87)
//  * codes -> UTF-8 of codes_to_chars     (REF/*/utils.py:18-37, REF/pretraining-data/converter.py:17-37)
#pragma once
#include "common.cuh"

namespace mimi {

// y[m] = sum_n x[n] * h[c + m*M - n*L]  (zero-phase prototype h of odd length 2c+1 on the L*sr_in grid),
// m < out_len[b]; samples in [out_len[b], out_stride) are written as zeros so the result is the padded
// [B,1,N] batch directly. One thread per output sample; x re-reads are served by L1 (neighbouring outputs
// share all but M/L of their taps), taps sit in L1/L2.
__global__ void __launch_bounds__(256) resample_kernel(const float* __restrict__ x, long long in_stride,
                                                       const int* __restrict__ in_len,
                                                       const int* __restrict__ out_len,
                                                       const float* __restrict__ taps, int c, int L, int M,
                                                       float* __restrict__ y, long long out_stride) {
  const int b = blockIdx.y;
  const long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= out_stride) return;
  float acc = 0.f;
  if (m < out_len[b]) {
    const int n_in = in_len[b];
    const long long p = m * M;                                   // fine-grid position of output m
    // taps index i = c + p - n*L in [0, 2c]  <=>  (p - c)/L <= n <= (p + c)/L
    long long nlo = (p - c + L - 1) / L;                          // ceil for p-c >= 0
    if (p - c < 0) nlo = 0;
    long long nhi = (p + c) / L;
    if (nhi > n_in - 1) nhi = n_in - 1;
    const float* xb = x + (long long)b * in_stride;
    for (long long n = nlo; n <= nhi; ++n) acc = fmaf(__ldg(xb + n), __ldg(taps + (c + p - n * L)), acc);
  }
  y[(long long)b * out_stride + m] = acc;
}

// ---- polyphase form for small L*M (16 k, 48 k, 8 k, 32 k, 12 k -> 24 k) ----------------------------------------------
// With n = (q + v)*M + r and m = q*L + phi the sum above is
//     y[q*L + phi] = sum_r sum_v  xs_r[q + v] * g[r][v][phi],   xs_r[k] = x[k*M + r],  g[r][v][phi] = h[c + phi*M - (v*M + r)*L]
// i.e. L*M plain stride-1 FIRs over the M de-interleaved input phases. A CTA stages the de-interleaved window of kQB output
// blocks q in shared memory (coalesced loads, skewed by one word per 32 so that lanes 8 words apart hit different banks);
// thread t owns Q = 8 consecutive blocks q and all L phases: per (r, v) it loads ONE new input sample and L taps (warp
// broadcast) for L*Q FMAs out of a register sliding window -- 6 FMAs per shared-memory word at L = 3, so the FP32 pipe is
// the bound, not the LSU. Results go back through shared memory and leave as coalesced float4 rows.
// FP32 roofline, not HBM: 16 k -> 24 k spends 80 FMAs (68 on non-zero taps; zero taps pad each sub-filter to a multiple of 8) per
// output sample for 6.7 bytes of traffic, i.e. 24 FLOP/B against the B200's ~11 FLOP/B fp32 ridge.
namespace rsp {
constexpr int kQ = 8;                 // output blocks per thread
constexpr int kThreads = 128;
constexpr int kQB = kQ * kThreads;    // output blocks per CTA
__host__ __device__ inline int skew(int k) { return k + (k >> 5); }
// floats of shared memory: taps [M][V][4] + de-interleaved window M * skew(kQB + V + kQ)
__host__ __device__ inline size_t smem_floats(int L, int M, int V) {
  const size_t xs = (size_t)M * (size_t)(skew(kQB + V + kQ) + 1);
  const size_t out = (size_t)kQB * L;
  return (size_t)M * V * 4 + (xs > out ? xs : out);
}
}  // namespace rsp

template <int L, int M>
__global__ void __launch_bounds__(rsp::kThreads) resample_poly_kernel(const float* __restrict__ x, long long in_stride,
                                                                      const int* __restrict__ in_len,
                                                                      const int* __restrict__ out_len,
                                                                      const float* __restrict__ g,   // [M][V][4]
                                                                      int V, int Vh, float* __restrict__ y,
                                                                      long long out_stride) {
  static_assert(L >= 1 && L <= 4, "phases per block");
  extern __shared__ float sm[];
  float* sg = sm;                                    // taps
  float* sx = sm + (size_t)M * V * 4;                // window, later the output tile
  const int xs_stride = rsp::skew(rsp::kQB + V + rsp::kQ) + 1;
  const int b = blockIdx.y;
  const long long q0 = (long long)blockIdx.x * rsp::kQB;
  const int tid = threadIdx.x;
  const int n_in = in_len[b], n_out = out_len[b];
  const float* xb = x + (long long)b * in_stride;
  for (int i = tid; i < M * V * 4; i += rsp::kThreads) sg[i] = g[i];
  // window: input blocks k in [q0 - Vh, q0 - Vh + kQB + V), sample n = k*M + r; consecutive threads read consecutive n
  const long long n_base = (q0 - Vh) * M;
  const int span = (rsp::kQB + V) * M;
  for (int t = tid; t < span; t += rsp::kThreads) {
    const long long n = n_base + t;
    const int k = t / M, r = t - k * M;
    sx[r * xs_stride + rsp::skew(k)] = (n >= 0 && n < n_in) ? __ldg(xb + n) : 0.f;
  }
  __syncthreads();
  float acc[L][rsp::kQ];
#pragma unroll
  for (int p = 0; p < L; ++p)
#pragma unroll
    for (int i = 0; i < rsp::kQ; ++i) acc[p][i] = 0.f;
  const int kb = tid * rsp::kQ;                      // first window slot of this thread's block 0 at v = -Vh
#pragma unroll
  for (int r = 0; r < M; ++r) {
    const float* xr = sx + r * xs_stride;
    const float4* gr = reinterpret_cast<const float4*>(sg) + (size_t)r * V;
    float w[rsp::kQ];
#pragma unroll
    for (int i = 0; i < rsp::kQ - 1; ++i) w[i] = xr[rsp::skew(kb + i)];
    for (int v0 = 0; v0 < V; v0 += rsp::kQ) {        // V is a multiple of kQ (zero taps pad it)
#pragma unroll
      for (int s = 0; s < rsp::kQ; ++s) {
        w[(s + rsp::kQ - 1) % rsp::kQ] = xr[rsp::skew(kb + v0 + s + rsp::kQ - 1)];
        const float4 t4 = gr[v0 + s];
        const float tp[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int p = 0; p < L; ++p)
#pragma unroll
          for (int i = 0; i < rsp::kQ; ++i) acc[p][i] = fmaf(w[(s + i) % rsp::kQ], tp[p], acc[p][i]);
      }
    }
  }
  __syncthreads();                                   // everyone is done with the window: reuse it as the output tile
#pragma unroll
  for (int i = 0; i < rsp::kQ; ++i)
#pragma unroll
    for (int p = 0; p < L; ++p) sx[(kb + i) * L + p] = acc[p][i];
  __syncthreads();
  const long long m0 = q0 * L;
  float* yb = y + (long long)b * out_stride;
  const int tile = rsp::kQB * L;
  if ((out_stride & 3) == 0 && ((reinterpret_cast<uintptr_t>(yb) & 15) == 0)) {
    for (int t = tid * 4; t < tile; t += rsp::kThreads * 4) {
      const long long m = m0 + t;
      if (m + 3 < out_stride) {
        float4 o = *reinterpret_cast<const float4*>(sx + t);
        if (m + 3 >= n_out) {
          if (m >= n_out) o.x = 0.f;
          if (m + 1 >= n_out) o.y = 0.f;
          if (m + 2 >= n_out) o.z = 0.f;
          o.w = 0.f;
        }
        *reinterpret_cast<float4*>(yb + m) = o;
      } else {
        for (int e = 0; e < 4; ++e)
          if (m + e < out_stride) yb[m + e] = (m + e < n_out) ? sx[t + e] : 0.f;
      }
    }
  } else {
    for (int t = tid; t < tile; t += rsp::kThreads) {
      const long long m = m0 + t;
      if (m < out_stride) yb[m] = (m < n_out) ? sx[t] : 0.f;
    }
  }
}

// code point = offset + k*codebook_size + code, written as the 1..4-byte UTF-8 form, frame-major / codebook-minor.
// bytes_per_frame is fixed per call because the host checked that no codebook's code-point range straddles a UTF-8 length
// boundary (and the caller that every code lies in [0, codebook_size)).
struct Utf8Params {
  const long long* codes;      // [B][K][T]
  int B, K;
  long long T;
  const int* frames;           // device [B] frames to convert per item, or nullptr -> T
  unsigned offset;
  int codebook_size;
  int bytes_per_frame;
  unsigned char* out;          // [B][out_stride]
  long long out_stride;
};

// Block = 256 consecutive frames of one item. Thread t reads its frame's K codes (for every k the warp reads 32 consecutive
// int64 = one 256-byte run), builds the frame's bytes in shared memory, and the block then writes its frames' bytes -- one
// contiguous run of the output row -- as 16-byte vectors (the shared buffer starts at the same offset mod 16 as the global
// run, so head and tail are the only byte-wise stores). The first version had one thread per (frame, codebook) storing 3-4
// single bytes each: 0.12 of the copy bandwidth at large sizes.
constexpr int kUtf8Frames = 256;
__global__ void __launch_bounds__(kUtf8Frames) codes_to_utf8_kernel(const Utf8Params p) {
  extern __shared__ __align__(16) unsigned char u8s[];
  const int b = blockIdx.y;
  const long long t0 = (long long)blockIdx.x * kUtf8Frames;
  const long long nfr = p.frames ? p.frames[b] : p.T;
  if (t0 >= nfr) return;
  const int n = (int)min((long long)kUtf8Frames, nfr - t0);          // frames of this block
  unsigned char* gout = p.out + (long long)b * p.out_stride + t0 * p.bytes_per_frame;
  const int mis = (int)(reinterpret_cast<uintptr_t>(gout) & 15);      // shared copy starts at the same offset mod 16
  unsigned char* row = u8s + mis;
  const int t = threadIdx.x;
  if (t < n) {
    unsigned char* o = row + t * p.bytes_per_frame;
    const long long* cb = p.codes + (long long)b * p.K * p.T + t0 + t;
    for (int k = 0; k < p.K; ++k) {
      const unsigned cp = p.offset + (unsigned)k * (unsigned)p.codebook_size + (unsigned)cb[(long long)k * p.T];
      if (cp < 0x80u) {
        *o++ = (unsigned char)cp;
      } else if (cp < 0x800u) {
        *o++ = (unsigned char)(0xC0u | (cp >> 6));
        *o++ = (unsigned char)(0x80u | (cp & 0x3Fu));
      } else if (cp < 0x10000u) {
        *o++ = (unsigned char)(0xE0u | (cp >> 12));
        *o++ = (unsigned char)(0x80u | ((cp >> 6) & 0x3Fu));
        *o++ = (unsigned char)(0x80u | (cp & 0x3Fu));
      } else {
        *o++ = (unsigned char)(0xF0u | (cp >> 18));
        *o++ = (unsigned char)(0x80u | ((cp >> 12) & 0x3Fu));
        *o++ = (unsigned char)(0x80u | ((cp >> 6) & 0x3Fu));
        *o++ = (unsigned char)(0x80u | (cp & 0x3Fu));
      }
    }
  }
  __syncthreads();
  const int total = n * p.bytes_per_frame;
  const int head = min(total, (16 - mis) & 15);                         // bytes up to the first 16-byte boundary
  const int body = (total - head) / 16;                                 // whole vectors
  if (t < head) gout[t] = row[t];
  const uint4* src = reinterpret_cast<const uint4*>(row + head);        // (u8s + mis + head) is 16-byte aligned
  uint4* dst = reinterpret_cast<uint4*>(gout + head);
  for (int i = t; i < body; i += kUtf8Frames) dst[i] = src[i];
  const int done = head + body * 16;
  if (t < total - done) gout[done + t] = row[done + t];
}

// codes int64 -> uint16 (the `codes.astype(np.uint16)` of REF/yodas2-mimi/process_shard.py:519-523, before the D2H copy: a
// quarter of the bytes cross PCIe). Grid-stride, 4 codes per thread: 32 B in, 8 B out.
__global__ void __launch_bounds__(256) codes_pack_u16_kernel(const long long* __restrict__ codes, long long n,
                                                             unsigned short* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x * 4;
  for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 3 < n && (reinterpret_cast<uintptr_t>(out + i) & 7) == 0 && (reinterpret_cast<uintptr_t>(codes + i) & 15) == 0) {
      const longlong2 a = *reinterpret_cast<const longlong2*>(codes + i);
      const longlong2 b = *reinterpret_cast<const longlong2*>(codes + i + 2);
      ushort4 o;
      o.x = (unsigned short)a.x; o.y = (unsigned short)a.y; o.z = (unsigned short)b.x; o.w = (unsigned short)b.y;
      *reinterpret_cast<ushort4*>(out + i) = o;
    } else {
      for (int e = 0; e < 4 && i + e < n; ++e) out[i + e] = (unsigned short)codes[i + e];
    }
  }
}

}  // namespace mimi
