// Encoder-transformer pieces that are not GEMMs: LayerNorm rows and causal sliding-window attention
// with RoPE applied while staging Q/K (MimiTransformerLayer, modeling_mimi.py:926-993; MimiSdpaAttention
// :852-916; rotary :515-611; mask = create_sliding_window_causal_mask, window 250, :1096-1102).
#pragma once
#include "common.cuh"

namespace mimi {

// nn.LayerNorm(512, eps=1e-5): one warp per row, two-pass statistics held in registers.
__global__ void __launch_bounds__(256) layernorm512_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                           const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           long long item_stride, const int* __restrict__ len,
                                                           int uniform_len, float* __restrict__ y_lo, int lob = 0) {
  // y_lo != nullptr: write the hi/lo TF32 split of the result into (y, y_lo) for a tensor-core consumer
  const int b = blockIdx.y;
  const int L = len ? len[b] : uniform_len;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x * 8 + warp;
  if (t >= L) return;
  const float* xr = x + (long long)b * item_stride + (long long)t * kHidden;
  float* yr = y + (long long)b * item_stride + (long long)t * kHidden;
  float4 v[4];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i] = ld_nc_f4(xr + (i * 32 + lane) * 4);
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.f / kHidden);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) * (1.f / kHidden) + 1e-5f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = (i * 32 + lane) * 4;
    const float4 g = ld_nc_f4(gamma + c), bt = ld_nc_f4(beta + c);
    float4 o;
    o.x = v[i].x * rstd * g.x + bt.x;
    o.y = v[i].y * rstd * g.y + bt.y;
    o.z = v[i].z * rstd * g.z + bt.z;
    o.w = v[i].w * rstd * g.w + bt.w;
    if (y_lo) store_split4_x(y, y_lo, (long long)b * item_stride + (long long)t * kHidden + c, o, lob);
    else *reinterpret_cast<float4*>(yr + c) = o;
  }
}

// Causal sliding-window attention, fp32. One CTA = (32-query tile, head, item); its whole key window
// (<= 32 + 249 keys) is staged in shared memory with RoPE applied on the way in, then each warp handles
// 4 queries: lanes own keys for the QK^T pass (conflict-free K rows of odd pitch) and own output dims
// for the PV pass.
//   qkv  [B][item_stride]  row t = [q(512) | k(512) | v(512)], head h at h*64 inside each block
//   out  [B][out_stride]   row t = 512 floats, head h at h*64
//   rope_cos/sin [pos][32] fp32 tables (cos/sin of pos * inv_freq[i]); position = row index
constexpr int kAttQT = 32;
constexpr int kAttKeys = kAttQT + kWindow - 1;     // 281
constexpr int kAttKPitch = kHeadDim + 1;           // 65: lane-varying rows hit distinct banks
constexpr size_t kAttSmemBytes =
    sizeof(float) * (size_t)(kAttKeys * kAttKPitch + kAttKeys * kHeadDim + kAttQT * kHeadDim + 8 * 256);

__global__ void __launch_bounds__(256) swa_attention_kernel(const float* __restrict__ qkv, long long item_stride,
                                                            float* __restrict__ out, long long out_stride,
                                                            const float* __restrict__ rope_cos,
                                                            const float* __restrict__ rope_sin,
                                                            const int* __restrict__ len, int uniform_len) {
  extern __shared__ __align__(16) float smem[];
  float* Vs = smem;                                  // [281][64]  (float4 stores: 16-byte aligned)
  float* Qs = Vs + kAttKeys * kHeadDim;              // [32][64]
  float* Ps = Qs + kAttQT * kHeadDim;                // [8 warps][256]
  float* Ks = Ps + 8 * 256;                          // [281][65]  (odd pitch, scalar access only)

  const int b = blockIdx.z, h = blockIdx.y;
  const int T = len ? len[b] : uniform_len;
  const int q0 = blockIdx.x * kAttQT;
  if (q0 >= T) return;
  const int q1 = min(q0 + kAttQT, T);                // exclusive
  const int klo = max(0, q0 - (kWindow - 1));
  const int nkeys = q1 - klo;                        // keys klo .. q1-1
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* base = qkv + (long long)b * item_stride;

  // stage K (rotated) and V: 16 float4 per 64-float row; pair (d, d+32) handled by one thread
  for (int idx = tid; idx < nkeys * 8; idx += 256) {
    const int r = idx >> 3, d4 = (idx & 7) * 4;      // d4 in {0,4,..,28}: covers d4..d4+3 and +32
    const int pos = klo + r;
    const float* row = base + (long long)pos * (3 * kHidden);
    const float4 k_lo = ld_nc_f4(row + kHidden + h * kHeadDim + d4);
    const float4 k_hi = ld_nc_f4(row + kHidden + h * kHeadDim + d4 + 32);
    const float4 c = ld_nc_f4(rope_cos + (long long)pos * 32 + d4);
    const float4 s = ld_nc_f4(rope_sin + (long long)pos * 32 + d4);
    float* kd = Ks + r * kAttKPitch + d4;
    // q*cos + rotate_half(q)*sin with rotate_half(u) = [-u[32:], u[:32]]  (modeling_mimi.py:580-611)
    kd[0] = k_lo.x * c.x - k_hi.x * s.x;  kd[32] = k_hi.x * c.x + k_lo.x * s.x;
    kd[1] = k_lo.y * c.y - k_hi.y * s.y;  kd[33] = k_hi.y * c.y + k_lo.y * s.y;
    kd[2] = k_lo.z * c.z - k_hi.z * s.z;  kd[34] = k_hi.z * c.z + k_lo.z * s.z;
    kd[3] = k_lo.w * c.w - k_hi.w * s.w;  kd[35] = k_hi.w * c.w + k_lo.w * s.w;
    const float4 v_lo = ld_nc_f4(row + 2 * kHidden + h * kHeadDim + d4);
    const float4 v_hi = ld_nc_f4(row + 2 * kHidden + h * kHeadDim + d4 + 32);
    *reinterpret_cast<float4*>(Vs + r * kHeadDim + d4) = v_lo;
    *reinterpret_cast<float4*>(Vs + r * kHeadDim + d4 + 32) = v_hi;
  }
  // stage Q (rotated, pre-scaled by 1/sqrt(64) = 0.125, exact power of two)
  for (int idx = tid; idx < (q1 - q0) * 8; idx += 256) {
    const int r = idx >> 3, d4 = (idx & 7) * 4;
    const int pos = q0 + r;
    const float* row = base + (long long)pos * (3 * kHidden);
    const float4 q_lo = ld_nc_f4(row + h * kHeadDim + d4);
    const float4 q_hi = ld_nc_f4(row + h * kHeadDim + d4 + 32);
    const float4 c = ld_nc_f4(rope_cos + (long long)pos * 32 + d4);
    const float4 s = ld_nc_f4(rope_sin + (long long)pos * 32 + d4);
    float* qd = Qs + r * kHeadDim + d4;
    qd[0] = (q_lo.x * c.x - q_hi.x * s.x) * 0.125f;  qd[32] = (q_hi.x * c.x + q_lo.x * s.x) * 0.125f;
    qd[1] = (q_lo.y * c.y - q_hi.y * s.y) * 0.125f;  qd[33] = (q_hi.y * c.y + q_lo.y * s.y) * 0.125f;
    qd[2] = (q_lo.z * c.z - q_hi.z * s.z) * 0.125f;  qd[34] = (q_hi.z * c.z + q_lo.z * s.z) * 0.125f;
    qd[3] = (q_lo.w * c.w - q_hi.w * s.w) * 0.125f;  qd[35] = (q_hi.w * c.w + q_lo.w * s.w) * 0.125f;
  }
  __syncthreads();

  float* P = Ps + warp * 256;
  for (int qi = warp; qi < q1 - q0; qi += 8) {
    const int i = q0 + qi;                           // query position
    const int jlo = max(0, i - (kWindow - 1));       // first visible key: j > i - 250
    const int n = i - jlo + 1;                       // visible keys jlo..i  (<= 250)
    const int r0 = jlo - klo;                        // smem row of key jlo
    const float* q = Qs + qi * kHeadDim;
    // scores: lane owns keys r0 + lane + 32*u
    float sc[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) sc[u] = 0.f;
#pragma unroll 4
    for (int d = 0; d < kHeadDim; ++d) {
      const float qd = q[d];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int jj = lane + 32 * u;
        const int r = min(r0 + jj, kAttKeys - 1);     // clamp keeps the read in bounds; masked below
        sc[u] = fmaf(qd, Ks[r * kAttKPitch + d], sc[u]);
      }
    }
    float m = -INFINITY;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (lane + 32 * u >= n) sc[u] = -INFINITY;
      m = fmaxf(m, sc[u]);
    }
    m = warp_max(m);
    float sum = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const float e = (lane + 32 * u < n) ? expf(sc[u] - m) : 0.f;
      sc[u] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int u = 0; u < 8; ++u) P[lane + 32 * u] = sc[u] * inv;
    __syncwarp();
    // out[d] = sum_j P[j] * V[j][d]; lane owns d = lane and lane + 32
    float o0 = 0.f, o1 = 0.f;
    for (int jj = 0; jj < n; ++jj) {
      const float pj = P[jj];
      const float* vr = Vs + (r0 + jj) * kHeadDim;
      o0 = fmaf(pj, vr[lane], o0);
      o1 = fmaf(pj, vr[lane + 32], o1);
    }
    float* orow = out + (long long)b * out_stride + (long long)i * kHidden + h * kHeadDim;
    orow[lane] = o0;
    orow[lane + 32] = o1;
    __syncwarp();
  }
}

}  // namespace mimi
