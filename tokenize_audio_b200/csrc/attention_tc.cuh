// Causal sliding-window attention on tcgen05 (MimiSdpaAttention, modeling_mimi.py:852-916; window 250, :1096-1102;
// rotary :515-611). Same arithmetic as the SIMT kernels in transformer.cuh -- fp32-equivalent scores and P*V through
// 3xTF32 (hi*hi + hi*lo + lo*hi), fp32 softmax -- but the two contractions run on the tensor cores:
//
//   unit   = (item, head, tile of 128 queries q0 .. q0+127); one persistent CTA per SM walks the units.
//   keys   = six chunks of 64: chunk c holds keys q0 - 256 + 64c .. +63 (the window of the tile's queries is
//            q0 - 249 .. q0 + 127); chunks that lie wholly before the item start or after the tile's last query
//            are skipped (the first tile of an item needs two chunks, not six).
//   S      : Q [128 x 64] (RoPE'd, pre-scaled by 1/8, hi/lo split, staged once per unit) times K_c^T [64 x 64] (RoPE'd,
//            split, staged by the threads through a 2-deep ring) -> TMEM columns [64c, 64c+64): all scores of the unit
//            (128 x 384 fp32) stay in tensor memory.
//   softmax: thread (row, slice) owns query row and 16 of every chunk's 64 columns. Pass A reads the scores back for the
//            masked row maximum; pass B, chunk by chunk, turns them into e = exp(s - max) (unnormalised, <= 1), splits
//            hi/lo and writes them as the A operand of the P*V MMA.
//   O      : P_c [128 x 64 keys] times V_c [64 keys x 64 d] (staged transposed, K-major) into one of two 64-column TMEM
//            buffers; every chunk's result is drained into fp32 registers (round-to-nearest adds, the tensor core
//            accumulates with truncation) while the next chunk runs. out = sum / row sum, split hi/lo for o_proj.
//   16 worker warps: warps w, w+4, w+8, w+12 share TMEM lane quarter w%4 (32 query rows) and take a quarter of the columns
//   each; K and V chunks are loaded into registers one chunk ahead of their shared-memory stores. A 17th warp only issues
//   MMAs: the workers signal "operands staged" on an mbarrier and never meet in a block-wide barrier inside a unit except to
//   exchange the row maxima and sums (when a worker warp issued the MMAs itself, every chunk waited for that warp).
#pragma once
#include "tc_gemm2.cuh"
#include "front_fused.cuh"

namespace mimi {
namespace atc {

constexpr int kWorkers = 512;                  // 16 worker warps; warp 16 issues the MMAs
constexpr int kThreads = kWorkers + 32;
constexpr int kSlices = 4;                     // column slices per TMEM lane quarter (warps w, w+4, w+8, w+12)
constexpr int kCols = 16;                      // score / output columns per thread and chunk
constexpr int kQT = 128;                       // queries per unit
constexpr int kKC = 64;                        // keys per chunk
constexpr int kChunks = 6;                     // key chunks per unit: keys q0 - 256 .. q0 + 127
constexpr int kQPanel = kQT * 128;             // 16 KB: 128 rows x 32 floats
constexpr int kKPanel = kKC * 128;             // 8 KB
constexpr int kQBytes = 4 * kQPanel;           // hi panel 0 | hi panel 1 | lo panel 0 | lo panel 1
constexpr int kKVBytes = 4 * kKPanel;
constexpr int kMaxLenB = 256;
constexpr int kSmem = 1024 + 2 * kQBytes + 2 * kKVBytes + 2 * kSlices * kQT * 4 + 256 + kMaxLenB * 4;
constexpr int kTmemCols = 512;                 // S: 384 columns, O: 2 x 64

struct Params {
  const float* qkv;            // [B][item_stride] rows of [q(512) | k(512) | v(512)]
  long long item_stride;
  float* out_hi;               // [B][out_stride] rows of 512, head h at h*64: TF32 hi / lo split of the result
  float* out_lo;
  long long out_stride;
  const float* rope_cos;       // [pos][32]
  const float* rope_sin;
  const int* len;              // device [B] rows per item or nullptr -> uniform_len
  int uniform_len;
  int B;
  int mt_max;                  // query tiles per item at the longest item
  int lo_bf16;                 // 1: out_lo is a bf16 array (mode 7)
  // Ragged call: the compact list of the query tiles that exist (item << 20 | tile, tile-major: the level-4 list of the GEMMs,
  // tc_gemm2.cuh). Walked from its END: later tiles see more key chunks (2, 4, 6, 6, ...), so the expensive units are dealt
  // first and the cheap ones level the tail; no unit past an item's end is ever visited (the mt_max x B grid walk left the SMs
  // idle for 31 % of a launch on a C2 batch, ncu round 2).
  const int* tiles;            // device [ntiles] or nullptr -> walk the mt_max x B grid
  int ntiles;
};

// exp(x) for x <= 0 through ex2.approx; the scaling by log2(e) is done in two pieces so that the argument carries no
// more than one rounding (relative error of the result ~3e-7 for the terms that matter, |x| < 8)
__device__ __forceinline__ float exp_neg(float x) {
  const float t = fmaf(x, 1.4426950216293335f, x * 1.9259629911266175e-8f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
  return e;
}

// Staging is split into a load half (global -> registers) and a store half (registers -> shared memory) so that the loads
// of the NEXT chunk are in flight while this chunk's MMAs are issued and waited for.
//
// q / k rows (RoPE applied, scaled) as a K-major SWIZZLE_128B hi/lo operand: [hi d 0-31 | hi d 32-63 | lo d 0-31 |
// lo d 32-63], each panel ROWS x 128 B. Task t -> row t/8 (position pos0 + row), dims 4g..4g+3 and 32+4g..32+4g+3
// (g = t%8: the two halves RoPE rotates into each other). Positions outside [0, T) are written as zeros.
struct RotRegs { float4 x0, x1, c, s; };

__device__ __forceinline__ void load_rot(RotRegs& r, int t, const float* __restrict__ src, const float* __restrict__ rope_cos,
                                         const float* __restrict__ rope_sin, int pos0, int T) {
  const int row = t >> 3, g = t & 7;
  const int pos = pos0 + row;
  r.x0 = r.x1 = r.c = r.s = make_float4(0.f, 0.f, 0.f, 0.f);       // zeros rotate to zeros
  if (pos >= 0 && pos < T) {
    const float* rp = src + (long long)pos * (3 * kHidden);
    r.x0 = ld_nc_f4(rp + g * 4); r.x1 = ld_nc_f4(rp + 32 + g * 4);
    r.c = ld_nc_f4(rope_cos + (long long)pos * 32 + g * 4); r.s = ld_nc_f4(rope_sin + (long long)pos * 32 + g * 4);
  }
}
template <int ROWS>
__device__ __forceinline__ void store_rot(const RotRegs& r, int t, uint32_t dst, float scale) {
  constexpr int PANEL = ROWS * 128;
  const int row = t >> 3, g = t & 7;
  float4 lo4, hi4;
  lo4.x = (r.x0.x * r.c.x - r.x1.x * r.s.x) * scale;  hi4.x = (r.x1.x * r.c.x + r.x0.x * r.s.x) * scale;
  lo4.y = (r.x0.y * r.c.y - r.x1.y * r.s.y) * scale;  hi4.y = (r.x1.y * r.c.y + r.x0.y * r.s.y) * scale;
  lo4.z = (r.x0.z * r.c.z - r.x1.z * r.s.z) * scale;  hi4.z = (r.x1.z * r.c.z + r.x0.z * r.s.z) * scale;
  lo4.w = (r.x0.w * r.c.w - r.x1.w * r.s.w) * scale;  hi4.w = (r.x1.w * r.c.w + r.x0.w * r.s.w) * scale;
  const uint32_t a = dst + (uint32_t)(row * 128 + ((g ^ (row & 7)) << 4));
  float4 h, l;
  split_tf32(lo4.x, h.x, l.x); split_tf32(lo4.y, h.y, l.y); split_tf32(lo4.z, h.z, l.z); split_tf32(lo4.w, h.w, l.w);
  sts128(a, h);                 sts128(a + 2 * PANEL, l);                 // dims g*4 .. +3      -> panel 0
  split_tf32(hi4.x, h.x, l.x); split_tf32(hi4.y, h.y, l.y); split_tf32(hi4.z, h.z, l.z); split_tf32(hi4.w, h.w, l.w);
  sts128(a + PANEL, h);         sts128(a + 3 * PANEL, l);                 // dims 32 + g*4 .. +3 -> panel 1
}

// 64 keys of v TRANSPOSED as the B operand of P*V: N = 64 head dims (rows), K = 64 keys; panel p holds keys 32p .. 32p+31
// as rows of 128 B. [hi p0 | hi p1 | lo p0 | lo p1]. Thread -> key tid%64, head dims 8*(tid/64) .. +7. Keys outside
// [0, T) are zeros (0 * NaN would not be).
struct VtRegs { float4 v0, v1; };

__device__ __forceinline__ void load_vt(VtRegs& r, const float* __restrict__ src, int pos0, int T) {
  const int key = threadIdx.x & 63, dg = threadIdx.x >> 6;
  const int pos = pos0 + key;
  r.v0 = r.v1 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pos >= 0 && pos < T) {
    const float* rp = src + (long long)pos * (3 * kHidden) + dg * 8;
    r.v0 = ld_nc_f4(rp); r.v1 = ld_nc_f4(rp + 4);
  }
}
__device__ __forceinline__ void store_vt(const VtRegs& r, uint32_t dst) {
  const int key = threadIdx.x & 63, dg = threadIdx.x >> 6;
  const uint32_t base = dst + (uint32_t)((key >> 5) * kKPanel + (key & 3) * 4);
  const int cq = (key & 31) >> 2;
  const float vv[8] = {r.v0.x, r.v0.y, r.v0.z, r.v0.w, r.v1.x, r.v1.y, r.v1.z, r.v1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int d = dg * 8 + j;
    float hi, lo;
    split_tf32(vv[j], hi, lo);
    const uint32_t a = base + (uint32_t)(d * 128 + ((cq ^ (d & 7)) << 4));
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(hi) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a + 2 * kKPanel), "f"(lo) : "memory");
  }
}

// D[128 x 64] (+)= A * B^T over K = 64 in 3xTF32: A = [hi p0 | hi p1 | lo p0 | lo p1] panels of 16 KB, B likewise with
// panels of 8 KB. One elected thread.
__device__ __forceinline__ void mma_3x(uint32_t tmem_d, uint32_t a_base, uint32_t b_base) {
  constexpr uint32_t idesc = tc::make_idesc(128, 64);
  const uint32_t da = tc::desc_lo(a_base), db = tc::desc_lo(b_base);
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t a_hi = da + ((p * kQPanel) >> 4) + 2 * k, a_lo = a_hi + ((2 * kQPanel) >> 4);
      const uint32_t b_hi = db + ((p * kKPanel) >> 4) + 2 * k, b_lo = b_hi + ((2 * kKPanel) >> 4);
      tc::umma_tf32_lo(tmem_d, a_hi, b_hi, idesc, (uint32_t)((p | k) != 0));
      tc::umma_tf32_lo(tmem_d, a_hi, b_lo, idesc, 1u);
      tc::umma_tf32_lo(tmem_d, a_lo, b_hi, idesc, 1u);
    }
}

__global__ void __launch_bounds__(kThreads, 1) swa_attention_tc_kernel(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t sQ = tc::smem_u32(smem);
  const uint32_t sP = sQ + kQBytes;
  const uint32_t sKV = sP + kQBytes;                                  // 2 stages of kKVBytes
  float* red = reinterpret_cast<float*>(smem + 2 * kQBytes + 2 * kKVBytes);   // {max, sum} x [4 slices][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * kSlices * kQT);
  uint64_t* bar_kv = bars;              // [2] the MMAs that read K/V stage s have completed
  uint64_t* bar_s = bars + 2;           // all scores of the unit are in TMEM
  uint64_t* bar_pv = bars + 3;          // [2] the P*V MMAs of a chunk of that parity have completed (P / O buffer of that parity)
  uint64_t* bar_ready = bars + 5;       // [2] operands of MMA job j are in shared memory (16 worker warps arrive), j % 2
  uint32_t* tmem_base_ptr = reinterpret_cast<uint32_t*>(bars + 7);
  int* s_len = reinterpret_cast<int*>(bars + 8);                      // [kMaxLenB] item lengths (a global load per unit otherwise)

  const int warp = tc::uniform_warp_idx(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::mbar_init(&bar_kv[0], 1); tc::mbar_init(&bar_kv[1], 1); tc::mbar_init(bar_s, 1); tc::mbar_init(&bar_pv[0], 1); tc::mbar_init(&bar_pv[1], 1);
    tc::mbar_init(&bar_ready[0], kWorkers / 32); tc::mbar_init(&bar_ready[1], kWorkers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const bool len_smem = p.len != nullptr && p.B <= kMaxLenB;
  if (len_smem)
    for (int i = threadIdx.x; i < p.B; i += kThreads) s_len[i] = __ldg(p.len + i);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(tmem_base_ptr)), "r"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_base_ptr;

  const int quarter = warp & 3, cs = warp >> 2;                       // TMEM lane quarter, column slice (16 of 64)
  const int row = quarter * 32 + lane;                                // query row of the tile owned by this thread
  const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
  const uint32_t tS = tmem_base + lane_off + (uint32_t)(cs * kCols);  // + 64c: this thread's columns of chunk c
  const uint32_t tO = tmem_base + lane_off + 384u + (uint32_t)(cs * kCols);   // + 64 * buffer
  const int units = (p.tiles ? p.ntiles : p.mt_max * p.B) * kHeads;
  uint32_t kvc = 0, pvc = 0, uc = 0, job = 0;                         // K/V stage uses, P*V chunks, units, MMA jobs so far

  // unit u -> geometry; false when the tile lies past the item's end (same decision in every role)
  auto decode = [&](int u, int& b, int& h, int& q0, int& T, int& c_lo, int& c_hi) {
    h = u % kHeads;
    const int t = u / kHeads;
    if (p.tiles) {
      const int e = __ldg(p.tiles + (p.ntiles - 1 - t));
      b = e >> 20;
      q0 = (e & 0xFFFFF) * kQT;
    } else {
      b = t % p.B;
      q0 = (t / p.B) * kQT;
    }
    T = len_smem ? s_len[b] : p.len ? __ldg(p.len + b) : p.uniform_len;
    if (q0 >= T) return false;
    const int kbase = q0 - 256;
    const int q_last = min(q0 + kQT, T) - 1;
    const int first_key = max(0, q0 - (kWindow - 1));
    c_lo = max(0, (first_key - kbase) / kKC);                         // first chunk with a visible key
    c_hi = (q_last - kbase) / kKC;                                    // last one (4 or 5)
    return true;
  };

  if (warp == kWorkers / 32) {
    // ---- MMA warp: one job per staged chunk (scores, then P*V), in the workers' order ----------------------------------
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int b, h, q0, T, c_lo, c_hi;
      if (!decode(u, b, h, q0, T, c_lo, c_hi)) continue;
      for (int pass = 0; pass < 2; ++pass)
        for (int c = c_lo; c <= c_hi; ++c, ++kvc, ++job) {
          const uint32_t st = kvc & 1u, pb = pvc & 1u;
          tc::mbar_wait(&bar_ready[job & 1u], (job >> 1) & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (tc::elect_one()) {
            if (pass == 0) {
              mma_3x(tmem_base + (uint32_t)(c * kKC), sQ, sKV + st * kKVBytes);
              tc::umma_commit(&bar_kv[st]);
              if (c == c_hi) tc::umma_commit(bar_s);
            } else {
              mma_3x(tmem_base + 384u + pb * 64u, pb ? sQ : sP, sKV + st * kKVBytes);
              tc::umma_commit(&bar_kv[st]);
              tc::umma_commit(&bar_pv[pb]);
            }
          }
          __syncwarp();
          if (pass == 1) ++pvc;
        }
    }
  } else
  for (int u = blockIdx.x; u < units; u += gridDim.x) {
    int b, h, q0, T, c_lo, c_hi;
    if (!decode(u, b, h, q0, T, c_lo, c_hi)) continue;
    const int kbase = q0 - 256;
    const float* item = p.qkv + (long long)b * p.item_stride + h * kHeadDim;
    // operands of the next MMA job are written: make them visible to the tensor core and tell the MMA warp
    auto signal_ready = [&]() {
      f0::fence_async_smem();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_ready[job & 1u]);
      ++job;
    };
    auto workers_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kWorkers) : "memory"); };

    // ---- Q, then the scores chunk by chunk; the row maximum follows two chunks behind the MMAs ----------------------------
    const int qi = q0 + row;                                          // this thread's query position
    // key of column j of chunk c (this thread's slice): kbase + 64c + 16cs + j; visible iff qi - 250 < key <= qi, key >= 0
    float mx = -INFINITY;
    auto scan_max = [&](int c) {
      uint32_t r[kCols];
      tc2::tmem_ld16_nowait(tS + (uint32_t)(c * kKC), r);
      tc2::tmem_ld_wait();
      const int k0 = kbase + c * kKC + cs * kCols;
#pragma unroll
      for (int j = 0; j < kCols; ++j) {
        const int key = k0 + j;
        const bool vis = key >= 0 && key <= qi && key > qi - kWindow;
        mx = fmaxf(mx, vis ? __uint_as_float(r[j]) : -INFINITY);
      }
    };
    RotRegs kr;
    {
      RotRegs q0r, q1r;
      load_rot(q0r, threadIdx.x, item, p.rope_cos, p.rope_sin, q0, T);
      load_rot(q1r, threadIdx.x + kWorkers, item, p.rope_cos, p.rope_sin, q0, T);
      load_rot(kr, threadIdx.x, item + kHidden, p.rope_cos, p.rope_sin, kbase + c_lo * kKC, T);
      store_rot<kQT>(q0r, threadIdx.x, sQ, 0.125f);
      store_rot<kQT>(q1r, threadIdx.x + kWorkers, sQ, 0.125f);
    }
    VtRegs vr;
    for (int c = c_lo; c <= c_hi; ++c, ++kvc) {
      const uint32_t st = kvc & 1u;
      tc::mbar_wait(&bar_kv[st], ((kvc >> 1) & 1u) ^ 1u);             // the MMAs of this stage's previous use are done
      store_rot<kKC>(kr, threadIdx.x, sKV + st * kKVBytes, 1.f);
      signal_ready();
      // next chunk's loads AFTER the proxy fence (the fence waits for every load in flight)
      if (c < c_hi) load_rot(kr, threadIdx.x, item + kHidden, p.rope_cos, p.rope_sin, kbase + (c + 1) * kKC, T);
      else load_vt(vr, item + 2 * kHidden, kbase + c_lo * kKC, T);
      if (c - 2 >= c_lo) {                                            // that previous use was chunk c-2 of this unit
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        scan_max(c - 2);
      }
    }
    tc::mbar_wait(bar_s, uc & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (c_hi - 1 >= c_lo) scan_max(c_hi - 1);
    scan_max(c_hi);
    red[cs * kQT + row] = mx;
    workers_sync();
    mx = fmaxf(fmaxf(red[row], red[kQT + row]), fmaxf(red[2 * kQT + row], red[3 * kQT + row]));   // finite for every query < T

    // ---- e = exp(s - max) and P*V, chunk by chunk. P alternates between its own buffer and the (now dead) Q buffer and O
    //      between two TMEM buffers, so the MMAs of chunk i run while the threads prepare chunk i+1 ---------------------------
    float oacc[kCols];
#pragma unroll
    for (int i = 0; i < kCols; ++i) oacc[i] = 0.f;
    float lsum = 0.f;
    const uint32_t pv0 = pvc;                                         // P*V chunk counter at the start of this unit
    for (int c = c_lo; c <= c_hi; ++c, ++kvc, ++pvc) {
      const uint32_t st = kvc & 1u, pb = pvc & 1u;
      tc::mbar_wait(&bar_kv[st], ((kvc >> 1) & 1u) ^ 1u);
      store_vt(vr, sKV + st * kKVBytes);
      float e[kCols];
      {
        uint32_t r[kCols];
        tc2::tmem_ld16_nowait(tS + (uint32_t)(c * kKC), r);
        tc2::tmem_ld_wait();
        const int k0 = kbase + c * kKC + cs * kCols;
#pragma unroll
        for (int j = 0; j < kCols; ++j) {
          const int key = k0 + j;
          const bool vis = key >= 0 && key <= qi && key > qi - kWindow;
          e[j] = vis ? exp_neg(__uint_as_float(r[j]) - mx) : 0.f;
          lsum += e[j];
        }
      }
      if (pvc >= pv0 + 2) {
        // chunk i-2 used the same P and O buffers: its MMAs are done, drain its result before both are reused
        tc::mbar_wait(&bar_pv[pb], ((pvc >> 1) - 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc2::drain_add<kCols>(tO + pb * 64u, oacc);
      }
      const uint32_t pbuf = pb ? sQ : sP;
      {
        // keys 16cs .. 16cs+15 of the chunk: panel cs/2, 16-byte pieces 4*(cs%2) .. +3 of this thread's row
        const uint32_t a = pbuf + (uint32_t)((cs >> 1) * kQPanel + row * 128);
        const int key7 = row & 7;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float4 h4, l4;
          split_tf32(e[4 * q], h4.x, l4.x); split_tf32(e[4 * q + 1], h4.y, l4.y);
          split_tf32(e[4 * q + 2], h4.z, l4.z); split_tf32(e[4 * q + 3], h4.w, l4.w);
          const uint32_t o = (uint32_t)((((cs & 1) * 4 + q) ^ key7) << 4);
          sts128(a + o, h4);
          sts128(a + o + 2 * kQPanel, l4);
        }
      }
      signal_ready();
      if (c < c_hi) load_vt(vr, item + 2 * kHidden, kbase + (c + 1) * kKC, T);
    }
    // the last two chunks (one, if the unit had a single chunk) are still in their O buffers
    for (uint32_t i = (pvc - pv0 >= 2 ? pvc - 2 : pvc - 1); i < pvc; ++i) {
      tc::mbar_wait(&bar_pv[i & 1u], (i >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      tc2::drain_add<kCols>(tO + (i & 1u) * 64u, oacc);
    }

    // ---- normalise, split, store (thread = one row, 16 consecutive head dims = one 64-byte segment per array) -------
    red[kSlices * kQT + cs * kQT + row] = lsum;
    workers_sync();
    if (qi < T) {
      const float* rs = red + kSlices * kQT + row;
      const float inv = 1.f / ((rs[0] + rs[kQT]) + (rs[2 * kQT] + rs[3 * kQT]));
      const long long o = (long long)b * p.out_stride + (long long)qi * kHidden + h * kHeadDim + cs * kCols;
#pragma unroll
      for (int q = 0; q < kCols / 4; ++q)
        store_split4_x(p.out_hi, p.out_lo, o + 4 * q,
                         make_float4(oacc[4 * q] * inv, oacc[4 * q + 1] * inv, oacc[4 * q + 2] * inv, oacc[4 * q + 3] * inv), p.lo_bf16);
    }
    ++uc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
  }
}

}  // namespace atc
}  // namespace mimi
