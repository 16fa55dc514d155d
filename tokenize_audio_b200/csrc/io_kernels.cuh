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

// One thread per (item, frame, codebook): code point = offset + k*codebook_size + code, written as the
// 1..4-byte UTF-8 form. byte_off[k] / bytes_per_frame are fixed per call because the host checked that no
// codebook's code-point range straddles a UTF-8 length boundary.
struct Utf8Params {
  const long long* codes;      // [B][K][T]
  int B, K;
  long long T;
  const int* frames;           // device [B] frames to convert per item, or nullptr -> T
  unsigned offset;
  int codebook_size;
  int bytes_per_frame;
  unsigned char byte_off[32];  // byte offset of codebook k inside a frame
  unsigned char* out;          // [B][out_stride]
  long long out_stride;
};

__global__ void __launch_bounds__(256) codes_to_utf8_kernel(const Utf8Params p) {
  const int b = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over T*K, frame-major
  const long long nfr = p.frames ? p.frames[b] : p.T;
  if (i >= nfr * p.K) return;
  const long long t = i / p.K;
  const int k = (int)(i - t * p.K);
  const unsigned cp = p.offset + (unsigned)k * (unsigned)p.codebook_size +
                      (unsigned)p.codes[((long long)b * p.K + k) * p.T + t];
  unsigned char* o = p.out + (long long)b * p.out_stride + t * p.bytes_per_frame + p.byte_off[k];
  if (cp < 0x80u) {
    o[0] = (unsigned char)cp;
  } else if (cp < 0x800u) {
    o[0] = (unsigned char)(0xC0u | (cp >> 6));
    o[1] = (unsigned char)(0x80u | (cp & 0x3Fu));
  } else if (cp < 0x10000u) {
    o[0] = (unsigned char)(0xE0u | (cp >> 12));
    o[1] = (unsigned char)(0x80u | ((cp >> 6) & 0x3Fu));
    o[2] = (unsigned char)(0x80u | (cp & 0x3Fu));
  } else {
    o[0] = (unsigned char)(0xF0u | (cp >> 18));
    o[1] = (unsigned char)(0x80u | ((cp >> 12) & 0x3Fu));
    o[2] = (unsigned char)(0x80u | ((cp >> 6) & 0x3Fu));
    o[3] = (unsigned char)(0x80u | (cp & 0x3Fu));
  }
}

}  // namespace mimi
