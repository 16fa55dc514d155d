// mimi_b200: C-ABI engine for the Mimi encode hot path on B200 (sm_100a). See include/mimi_b200.h for the
// boundary and DESIGN.md for the layout / kernel notes. No CPU fallback anywhere in this file.
#include "../../include/mimi_b200.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <emmintrin.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "common.cuh"
#include "gemm_simt.cuh"
#include "io_kernels.cuh"
#include "rvq.cuh"
#include "tc_gemm.cuh"
#include "tc_gemm2.cuh"
#include "front_fused.cuh"
#include "front_f16.cuh"
#include "rvq_tc.cuh"
#include "rvq_f16.cuh"
#include "tc_gemm5.cuh"
#include "tc_gemm7.cuh"
#include "transformer.cuh"
#include "attention_tc.cuh"
#include "decode_kernels.cuh"

using namespace mimi;

namespace {

struct ConvGeom { int cin, cout, k, stride; };
// SEANet convs in execution order (MimiEncoder.__init__, modeling_mimi.py:454-488)
const ConvGeom kConv[MIMI_B200_NUM_CONVS] = {
    {1, 64, 7, 1},    {64, 32, 3, 1},   {32, 64, 1, 1},   {64, 128, 8, 4},   {128, 64, 3, 1},
    {64, 128, 1, 1},  {128, 256, 10, 5}, {256, 128, 3, 1}, {128, 256, 1, 1},  {256, 512, 12, 6},
    {512, 256, 3, 1}, {256, 512, 1, 1}, {512, 1024, 16, 8}, {1024, 512, 3, 1}};
const int kLevelStride[5] = {4, 5, 6, 8, 2};   // level l+1 = ceil(level l / stride)
constexpr int kRopeMaxPos = 65536;             // 25 Hz positions (43.7 min of audio per item); reference chunks are <= 60 s
constexpr int kStageSlots = 4;

struct LayerDev {
  float *ln1_w, *ln1_b, *qkv_wt, *o_wt, *ls1, *ln2_w, *ln2_b, *fc1_wt, *fc2_wt, *ls2;
};

// float offsets of every activation buffer inside the caller's workspace (channels-last, dense
// [B][rows_max(level)][C])
struct Plan {
  int B = 0, K = 0;
  long long N = 0;
  int rows[6] = {0, 0, 0, 0, 0, 0};           // max rows per level: N, /4, /20, /120, T25, T
  long long a0, r1, d1, r2, d2, r3, d3, r4, d4, z, y, qkv, att, ffn, e, rp;   // float offsets
  long long ints;                              // byte offset of the int region (lens[6][B], prefix[B+1])
  size_t bytes = 0;
};

struct TapInfo { long long off; int level; int C; };

// ---- tensor-core (tcgen05 3xTF32) path ---------------------------------------------------------------
// A weight matrix [N][K] (K-major) in the operand formats of the two tensor-core generations, with its TMA maps
struct TcWeight {
  // mode 7: TF32 hi / lo (fp32, SWIZZLE_128B boxes of 32 floats) and bf16(hi) (SWIZZLE_64B)
  float* hi = nullptr;
  float* lo = nullptr;
  uint16_t* hib = nullptr;
  CUtensorMap map_hi, map_lo;              // box 32 x BN rows (the fused front end's resident W1 / W2)
  CUtensorMap m_hi[3], m_lo[3], m_hib[3];  // box rows 128, 64, 32 (the pair GEMM stages bnp / 2 rows per CTA)
  // mode 9: rows scaled by 2^e_n so that max |W'[n,:]| lies in [2^13, 2^14); h16 = fp16(W'), l16 = fp16(W' - h16),
  // s16 = fp16(h16 / 2048) (the factor of the activations' scaled lo parts); wscale[n] = 2^-e_n undoes the scaling in the epilogue.
  // One array [3][N][K] (hi | lo | hs); SWIZZLE_64B boxes of 32 x {128, 64, 32} rows.
  uint16_t* f16 = nullptr;
  // per-column affine of the epilogue, out = act(acc * cmul + cadd): cmul[0] (mode 7) = LayerScale or 1, cmul[1] (mode 9) = that
  // times the weight unscaling 2^-e_n; cadd = bias * LayerScale or 0
  float* cmul[2] = {nullptr, nullptr};
  float* cadd = nullptr;
  CUtensorMap map_f16[3][3];               // [hi, lo, hs][box rows 128, 64, 32]
  int N = 0, K = 0, BN = 0;
  // host copies of the mode-9 pack (R1a / R1b only: the fused front end keeps them resident in its own shared-memory layout)
  std::vector<uint16_t> h_f16;
  std::vector<float> h_cmul9, h_cadd;
};
// hi/lo activation pair, channels-last with `front` zero halo rows before row 0 of every item
// (mode 7: hi fp32, lo bf16; mode 9: both fp16 -- item_stride counts ELEMENTS per item, hi / lo are float offsets of the arrays)
struct SplitBuf { long long hi = 0, lo = 0, item_stride = 0; int front = 0, back = 0, C = 0, level = 0; };
constexpr int kHalo = 8;     // >= max(k - stride) = 8 (front) and >= max(stride) - 1 = 7 (back)

struct PlanTC {
  int B = 0, K = 0;
  long long N = 0;
  int rows[6] = {0, 0, 0, 0, 0, 0};
  long long d1, d2, d3, z, qkv, e, rp;                         // raw fp32 buffers (float offsets)
  SplitBuf s_h1, s_d1, s_r2, s_h2, s_d2, s_r3, s_h3, s_d3, s_r4, s_h4, s_d4, s_y, s_att, s_ffn, s_zp, s_e;
  long long ints;
  size_t bytes = 0;
};

struct MapSet { std::vector<CUtensorMap> maps; uint64_t built = 0; };
struct MapKey {
  const void* ws; int B; long long N; int layout;     // layout: the kernel generation whose workspace plan the offsets come from
  bool operator<(const MapKey& o) const {
    if (ws != o.ws) return ws < o.ws;
    if (B != o.B) return B < o.B;
    if (N != o.N) return N < o.N;
    return layout < o.layout;
  }
};

}  // namespace

struct mimi_b200 {
  int device = -1;
  std::string err;
  bool loaded = false;
  long long launches = 0;
  // weights (device)
  float* conv_wt[MIMI_B200_NUM_CONVS] = {};    // [K = k*Cin][Cout]; conv 0 keeps [64][7]
  float* conv_b[MIMI_B200_NUM_CONVS] = {};
  LayerDev layer[MIMI_B200_NUM_LAYERS] = {};
  float* down_wt = nullptr;                    // [2048][512]
  float* proj_wt = nullptr;                    // [512][512]: cols 0..255 semantic, 256..511 acoustic
  float* embed = nullptr;                      // [32][2048][256]
  float* embed_t = nullptr;                    // [32][256][2048]
  float* enorm = nullptr;                      // [32][2048]
  float* rope_cos = nullptr;                   // [kRopeMaxPos][32]
  float* rope_sin = nullptr;
  std::vector<void*> allocs;
  // pinned staging ring for small host->device int arrays
  int* stage[kStageSlots] = {};
  size_t stage_cap = 0;                        // ints per slot
  cudaEvent_t stage_ev[kStageSlots] = {};
  int stage_next = 0;
  // scratch for resample / utf8 length arrays (device)
  int* dev_ints = nullptr;
  size_t dev_ints_cap = 0;
  int* range_flag = nullptr;                   // pinned host copy of g_f16_overflow, refreshed behind every mode-9 encode
  cudaEvent_t dev_ints_ev = nullptr;           // recorded behind the last kernel that reads dev_ints: the next call's upload
                                               // (possibly on another stream) waits for it before overwriting the table
  // resampler taps cache: key (sr_in << 32 | sr_out)
  struct Taps { float* d; int c, L, M; float* g; int V, Vh; };   // g: polyphase table [M][V][4] (resample_poly_kernel) or nullptr
  std::map<unsigned long long, Taps> taps;
  // debug knobs and last plan
  int dbg_layers = MIMI_B200_NUM_LAYERS;
  int dbg_last_conv = MIMI_B200_NUM_CONVS - 1;
  // optional per-launch CUDA-event profile (debug_set key 2): events[i] closes launch prof_ids[i]
  bool sync_debug = false;             // MIMI_B200_SYNC=1: synchronise after every launch, report the first failure
  std::string sync_err;
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_pool;
  std::vector<int> prof_ids;           // -1 = "begin" marker
  size_t prof_n = 0;
  Plan last;
  void* last_ws = nullptr;
  // tensor-core path
  int mode = 9;                                // 9 = fused front end + CTA-pair tcgen05 GEMM with fp16 hi/lo operands (3 kind::f16 passes) +
                                               //     tcgen05 attention + tensor-core RVQ (default);
                                               // 7 = the same with TF32 hi (fp32) + bf16 lo operands: fp32 range, the fallback of mode 9;
                                               // 0 = every layer on fp32 FFMA (exact baseline for accuracy bisection)
  int last_mode = 0;
  int exp_chunk_kb = 0;                        // k-blocks per accumulation chunk (debug_set key 5; 0 = default 4)
  f0::Consts f0_consts;
  f1::Consts f1_consts;                        // the fp16 front end (front_f16.cuh): L0 weights, affines of R1a / R1b
  uint4* f1_wimg = nullptr;                    //   and the shared-memory image of W1 / W2 (hi | lo | hs, swizzle applied)
  int exp_front_tf32 = 0;                      // debug_set key 17: mode 9 runs round 1's front end (TF32 internals, front_fused.cuh)
  int exp_gelu_erff = 0;                       // debug_set key 21: fc1's GELU through erff (25 instructions) instead of gelu_fast
  int exp_no_taps = 0;                         // debug_set key 20: convs never run as tap groups (tc_gemm7.cuh), i.e. round 2's schedule
  int exp_att_grid = 0;                        // debug_set key 18: attention walks the mt_max x B grid (round 1's schedule)
  std::vector<int> len0_host;                  // samples per item of the call in flight (the front end's tile count)
  int num_sms = 148;
  long long item_tiles[6] = {0, 0, 0, 0, 0, 0};   // sum over items of ceil(rows_at_level / 128) for the call in flight
  const int* tile_ptr[6] = {};                 // ragged call in flight: compact 128-row tile lists per level (device), or nullptr
  int tile_cnt[6] = {};
  int exp_resample_simple = 0;                 // debug_set key 16: first-draft one-thread-per-output resampler (A/B of resample_poly_kernel)
  int exp_no_tile_list = 0;                    // debug_set key 13: walk the mt_max x B grid and skip (the old schedule)
  int phase = 0, front_b0 = 0, front_b1 = 0;   // mimi_b200_encode_phase: which part of the pipeline the call in flight runs
  int exp_linear_k = 0;                        // debug_set key 11: k-blocks in linear order (no tap grouping)
  int exp_no_flat = 0;                         // debug_set key 10: never flatten the linears' row dimension across items
  int num_clusters = 74;                       // co-resident CTA pairs of the cta_group::2 GEMM (tc_gemm5.cuh)
  int exp_pair_n128 = 0;                       // debug_set key 9: 1 = 256-column pair tiles wherever N allows (default: only for K > 2048)
  PFN_cuTensorMapEncodeTiled_v12000 encode_tiled = nullptr;
  TcWeight tc_conv[MIMI_B200_NUM_CONVS];       // convs 1..13 (conv 0 is a direct SIMT conv)
  TcWeight tc_qkv[MIMI_B200_NUM_LAYERS], tc_o[MIMI_B200_NUM_LAYERS], tc_fc1[MIMI_B200_NUM_LAYERS], tc_fc2[MIMI_B200_NUM_LAYERS];
  TcWeight tc_down, tc_proj;
  float *embed_hi = nullptr, *embed_lo = nullptr;   // TF32 split of the materialised codebooks (rvq_tc.cuh)
  CUtensorMap map_embed_hi, map_embed_lo;
  uint16_t* embed16 = nullptr;                      // fp16 pair of the row-scaled codebooks: [hi | lo][32 * 2048][256] (rvq_f16.cuh)
  float* embed_m2s = nullptr;                       // [32][2048] -2 * 2^-s of each code's row scaling
  CUtensorMap map_embed16_hi, map_embed16_lo;
  int exp_rvq_tf32 = 0;                             // debug_set key 19: mode 9 runs the TF32 RVQ (rvq_tc.cuh)
  std::map<MapKey, MapSet> amap_cache;   // activation maps per (workspace, B, N)
  PlanTC last_tc;
  bool last_was_tc = false;
  // decode direction (decode_host.inl): fp32 weights in the [K][N] layout of the FFMA GEMM
  bool dec_loaded = false;
  std::vector<void*> dec_allocs;
  float *dec_proj_wt = nullptr, *dec_up_w = nullptr, *dec_in_wt = nullptr, *dec_in_b = nullptr, *dec_out_wt = nullptr, *dec_out_b = nullptr;
  float *dec_up_wt[4] = {}, *dec_up_b[4] = {}, *dec_ra_wt[4] = {}, *dec_ra_b[4] = {}, *dec_rb_wt[4] = {}, *dec_rb_b[4] = {};
  LayerDev dlayer[MIMI_B200_NUM_LAYERS] = {};
  int* dec_bad = nullptr;
};

static std::string g_create_err;

// create / load_weights / destroy switch to the handle's device and restore the caller's current device on return
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); } }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define CUDA_TRY(h, expr)                                                                     \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e__) + " " + (h)->sync_err;   \
      return MIMI_B200_ERR_CUDA;                                                              \
    }                                                                                         \
  } while (0)

static int fail(mimi_b200* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_err = msg;
  return code;
}

// record "launch `id` just ended" on the stream (profiling only)
static void mark(mimi_b200* h, int id, cudaStream_t st) {
  if (h->sync_debug && id >= 0 && h->sync_err.empty()) {
    static const bool trace = getenv("MIMI_B200_TRACE") != nullptr;    // with MIMI_B200_SYNC: name every launch as it is awaited
    if (trace) { fprintf(stderr, "[mimi_b200] waiting for launch kind %d (#%lld)\n", id, (long long)h->launches); fflush(stderr); }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) h->sync_err = "launch kind " + std::to_string(id) + " (#" + std::to_string(h->launches) + "): " + cudaGetErrorString(e);
  }
  if (!h->prof_on) return;
  if (h->prof_n == h->prof_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    h->prof_pool.push_back(e);
    h->prof_ids.push_back(0);
  }
  h->prof_ids[h->prof_n] = id;
  cudaEventRecord(h->prof_pool[h->prof_n], st);
  h->prof_n++;
}

static int dev_upload(mimi_b200* h, float** dst, const std::vector<float>& src) {
  void* p = nullptr;
  CUDA_TRY(h, cudaMalloc(&p, src.size() * sizeof(float)));
  h->allocs.push_back(p);
  CUDA_TRY(h, cudaMemcpy(p, src.data(), src.size() * sizeof(float), cudaMemcpyHostToDevice));
  *dst = static_cast<float*>(p);
  return MIMI_B200_OK;
}

// [N][K] (out, in) -> [K][N]
static std::vector<float> transpose_nk(const float* w, int N, int K) {
  std::vector<float> t((size_t)N * K);
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) t[(size_t)k * N + n] = w[(size_t)n * K + k];
  return t;
}

// conv weight [Cout][Cin][k] -> Wt[(tau*Cin + ci)][Cout]
static std::vector<float> pack_conv(const float* w, int cout, int cin, int k) {
  std::vector<float> t((size_t)cout * cin * k);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int tau = 0; tau < k; ++tau)
        t[((size_t)tau * cin + ci) * cout + co] = w[((size_t)co * cin + ci) * k + tau];
  return t;
}

static void free_weights(mimi_b200* h) {
  for (void* p : h->allocs) cudaFree(p);
  h->allocs.clear();
  h->loaded = false;
}

// pinned staging slots for the small integer tables (lengths, tile lists) of a call; growing them costs a device
// synchronisation and four pinned allocations, so workspace_bytes() pre-sizes them for the batch it is asked about
static int ensure_stage(mimi_b200* h, size_t ints) {
  if (ints <= h->stage_cap) return MIMI_B200_OK;
  CUDA_TRY(h, cudaDeviceSynchronize());
  for (int i = 0; i < kStageSlots; ++i) {
    if (h->stage[i]) cudaFreeHost(h->stage[i]);
    h->stage[i] = nullptr;
  }
  h->stage_cap = std::max<size_t>(4096, ints + ints / 4);
  for (int i = 0; i < kStageSlots; ++i) CUDA_TRY(h, cudaMallocHost((void**)&h->stage[i], h->stage_cap * sizeof(int)));
  return MIMI_B200_OK;
}

static int stage_ints(mimi_b200* h, const std::vector<int>& v, int* d_dst, cudaStream_t st) {
  if (int rc = ensure_stage(h, v.size())) return rc;
  const int s = h->stage_next;
  h->stage_next = (s + 1) % kStageSlots;
  CUDA_TRY(h, cudaEventSynchronize(h->stage_ev[s]));
  std::memcpy(h->stage[s], v.data(), v.size() * sizeof(int));
  CUDA_TRY(h, cudaMemcpyAsync(d_dst, h->stage[s], v.size() * sizeof(int), cudaMemcpyHostToDevice, st));
  CUDA_TRY(h, cudaEventRecord(h->stage_ev[s], st));
  return MIMI_B200_OK;
}

static int ensure_dev_ints(mimi_b200* h, size_t n) {
  if (n <= h->dev_ints_cap) return MIMI_B200_OK;
  CUDA_TRY(h, cudaDeviceSynchronize());
  if (h->dev_ints) cudaFree(h->dev_ints);
  h->dev_ints_cap = std::max<size_t>(4096, n * 2);
  CUDA_TRY(h, cudaMalloc((void**)&h->dev_ints, h->dev_ints_cap * sizeof(int)));
  return MIMI_B200_OK;
}

static Plan make_plan(int B, long long N, int K) {
  Plan p;
  p.B = B; p.K = K; p.N = N;
  long long L = N;
  p.rows[0] = (int)L;
  for (int l = 0; l < 5; ++l) { L = (L + kLevelStride[l] - 1) / kLevelStride[l]; p.rows[l + 1] = (int)L; }
  long long off = 0;
  auto take = [&](int level, int C) {
    const long long o = off;
    long long n = (long long)B * p.rows[level] * C;
    n = (n + 63) / 64 * 64;                  // keep every buffer 256-byte aligned
    off += n;
    return o;
  };
  p.a0 = take(0, 64);  p.r1 = take(0, 32);
  p.d1 = take(1, 128); p.r2 = take(1, 64);
  p.d2 = take(2, 256); p.r3 = take(2, 128);
  p.d3 = take(3, 512); p.r4 = take(3, 256);
  p.d4 = take(4, 1024);
  p.z = take(4, 512);  p.y = take(4, 512);  p.qkv = take(4, 1536);  p.att = take(4, 512);  p.ffn = take(4, 2048);
  p.e = take(5, 512);  p.rp = take(5, 512);
  p.ints = off * (long long)sizeof(float);
  p.bytes = (size_t)p.ints + sizeof(int) * (size_t)(7 * B + 1 + 64);
  return p;
}

template <int BM, int BN, int TN>
static void launch_gemm_t(const GemmParams& p, int B, int lout_max, cudaStream_t st) {
  dim3 grid((lout_max + BM - 1) / BM, p.N / BN, B);
  gemm_f32_kernel<BM, BN, TN><<<grid, 256, 0, st>>>(p);
}

static int launch_gemm(mimi_b200* h, const GemmParams& p, int B, int max_len_in, cudaStream_t st, int prof_id) {
  const int lout_max = (max_len_in + p.stride - 1) / p.stride;
  if (lout_max <= 0 || B <= 0) return MIMI_B200_OK;
  if (p.K % 16 || p.Cin % 16 || p.N % 32) return fail(h, MIMI_B200_ERR_ARG, "gemm: unsupported shape");
  if (p.N % 128 == 0) launch_gemm_t<128, 128, 8>(p, B, lout_max, st);
  else if (p.N % 64 == 0) launch_gemm_t<128, 64, 4>(p, B, lout_max, st);
  else launch_gemm_t<256, 32, 4>(p, B, lout_max, st);
  h->launches++;
  mark(h, prof_id, st);
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}

// ---- resampler taps (same design as oracle/resample_oracle.py: Kaiser-windowed sinc) -----------------
static double bessel_i0(double x) {
  double sum = 1.0, term = 1.0;
  const double q = x * x / 4.0;
  for (int k = 1; k < 500; ++k) {
    term *= q / ((double)k * (double)k);
    sum += term;
    if (term < 1e-20 * sum) break;
  }
  return sum;
}

static void design_taps(int sr_in, int sr_out, std::vector<float>& taps, int& c, int& L, int& M) {
  long long a = sr_in, b = sr_out;
  while (b) { const long long t = a % b; a = b; b = t; }
  L = (int)(sr_out / a); M = (int)(sr_in / a);
  const double kZeros = 32.0, kRolloff = 0.945, kBeta = 14.769656459379492;
  const double fc = kRolloff * 0.5 / std::max(L, M);
  c = (int)std::ceil(kZeros / (2.0 * fc));
  const int n = 2 * c + 1;
  std::vector<double> hcoef(n);
  const double i0b = bessel_i0(kBeta);
  double sum = 0.0;
  for (int i = 0; i < n; ++i) {
    const double t = (double)(i - c);
    const double xx = 2.0 * fc * t;
    const double sinc = (t == 0.0) ? 1.0 : std::sin(M_PI * xx) / (M_PI * xx);
    const double r = (double)(i - c) / (double)c;
    const double w = bessel_i0(kBeta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
    hcoef[i] = 2.0 * fc * sinc * w;
    sum += hcoef[i];
  }
  taps.resize(n);
  for (int i = 0; i < n; ++i) taps[i] = (float)(hcoef[i] * (double)L / sum);
}

#include "tc_host.inl"
#include "decode_host.inl"

// ratios the register-tiled polyphase resampler is instantiated for (16 k, 48 k, 8 k, 32 k, 12 k, 96 k -> 24 k, identity)
static bool resample_poly_supported(int L, int M) {
  return (L == 3 && M == 2) || (L == 1 && M == 2) || (L == 3 && M == 1) || (L == 3 && M == 4) || (L == 2 && M == 1) ||
         (L == 1 && M == 4) || (L == 1 && M == 1);
}

static int utf8_len(unsigned cp) { return cp < 0x80u ? 1 : cp < 0x800u ? 2 : cp < 0x10000u ? 3 : 4; }

// =====================================================================================================
extern "C" {

int mimi_b200_abi_version(void) { return MIMI_B200_ABI_VERSION; }

const char* mimi_b200_last_error(const mimi_b200_t* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

int64_t mimi_b200_launch_count(const mimi_b200_t* h) { return h ? h->launches : 0; }

int mimi_b200_range_overflow(mimi_b200_t* h, int reset) {
  if (!h || !h->range_flag) return 0;
  const int v = *h->range_flag;
  if (reset) {
    DeviceGuard guard;
    cudaSetDevice(h->device);
    const int zero = 0;
    cudaDeviceSynchronize();
    cudaMemcpyToSymbol(g_f16_overflow, &zero, sizeof(int));
    *h->range_flag = 0;
  }
  return v;
}

int64_t mimi_b200_encoded_frames(int64_t n) {
  for (int l = 0; l < 5; ++l) n = (n + kLevelStride[l] - 1) / kLevelStride[l];
  return n;
}

int mimi_b200_create(mimi_b200_t** out, int device_ordinal) {
  if (!out) return fail(nullptr, MIMI_B200_ERR_ARG, "create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, MIMI_B200_ERR_CUDA,
                std::string("create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if (device_ordinal < 0 || device_ordinal >= n) return fail(nullptr, MIMI_B200_ERR_ARG, "create: bad device ordinal");
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device_ordinal)) != cudaSuccess)
    return fail(nullptr, MIMI_B200_ERR_CUDA, cudaGetErrorString(e));
  if (prop.major != 10)
    return fail(nullptr, MIMI_B200_ERR_CUDA, "create: this library is built for sm_100a (B200) only");
  DeviceGuard guard;
  mimi_b200* h = new mimi_b200();
  h->device = device_ordinal;
  h->num_sms = prop.multiProcessorCount;
  h->sync_debug = getenv("MIMI_B200_SYNC") != nullptr;
  if ((e = cudaSetDevice(device_ordinal)) != cudaSuccess) { delete h; return fail(nullptr, MIMI_B200_ERR_CUDA, cudaGetErrorString(e)); }
  for (int i = 0; i < kStageSlots; ++i) cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->dev_ints_ev, cudaEventDisableTiming);
  if (cudaMallocHost((void**)&h->range_flag, sizeof(int)) == cudaSuccess) *h->range_flag = 0;
  else { h->range_flag = nullptr; cudaGetLastError(); }
  cudaFuncSetAttribute(swa_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttSmemBytes);
  cudaFuncSetAttribute(rvq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRvqSmemBytes);
  cudaFuncSetAttribute(tcp::tcp_gemm_kernel<256, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::Cfg<256, 1>::SMEM);
  cudaFuncSetAttribute(tcp::tcp_gemm_kernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::Cfg<128, 1>::SMEM);
  cudaFuncSetAttribute(tcp::tcp_gemm_kernel<64, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::Cfg<64, 1>::SMEM);
  cudaFuncSetAttribute(tcp::tcp_gemm_kernel<256, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::Cfg<256, 3>::SMEM);
  cudaFuncSetAttribute(tcp::tcp_gemm_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::Cfg<128, 3>::SMEM);
  cudaFuncSetAttribute(tcp::tcp_gemm_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::Cfg<64, 3>::SMEM);
  {
    // how many CTA pairs of the widest instance fit at once (one per TPC unless the device says otherwise)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * (h->num_sms / 2)); cfg.blockDim = dim3(tcp::kThreads); cfg.dynamicSmemBytes = tcp::Cfg<256, 1>::SMEM;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int ncl = 0;
    if (cudaOccupancyMaxActiveClusters(&ncl, tcp::tcp_gemm_kernel<256, 1>, &cfg) == cudaSuccess && ncl > 0)
      h->num_clusters = std::min(ncl, h->num_sms / 2);
    else { cudaGetLastError(); h->num_clusters = h->num_sms / 2; }
  }
  cudaFuncSetAttribute(tcg::tcp_taps_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::Cfg<128, 2>::SMEM);
  cudaFuncSetAttribute(tcg::tcp_taps_kernel<128, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::Cfg<128, 3>::SMEM);
  cudaFuncSetAttribute(tcg::tcp_taps_kernel<64, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::Cfg<64, 3>::SMEM);
  cudaFuncSetAttribute(tcg::tcp_taps_kernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tcg::Cfg<64, 2>::SMEM);
  cudaFuncSetAttribute(atc::swa_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::kSmem);
  cudaFuncSetAttribute(rvqtc::rvq_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rvqtc::kSmem);
  cudaFuncSetAttribute(rvq16::rvq_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rvq16::kSmem);
  cudaFuncSetAttribute(f0::front_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f0::kSmem);
  cudaFuncSetAttribute(f1::front_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, f1::kSmem);
  cudaFuncSetAttribute(tc2::tc_shift_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  if ((e = cudaGetLastError()) != cudaSuccess) { delete h; return fail(nullptr, MIMI_B200_ERR_CUDA, cudaGetErrorString(e)); }
  *out = h;
  return MIMI_B200_OK;
}

void mimi_b200_destroy(mimi_b200_t* h) {
  if (!h) return;
  DeviceGuard guard;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  free_weights(h);
  for (int i = 0; i < kStageSlots; ++i) {
    if (h->stage[i]) cudaFreeHost(h->stage[i]);
    if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]);
  }
  if (h->dev_ints) cudaFree(h->dev_ints);
  if (h->dev_ints_ev) cudaEventDestroy(h->dev_ints_ev);
  if (h->range_flag) cudaFreeHost(h->range_flag);
  for (void* p : h->dec_allocs) cudaFree(p);
  if (h->dec_bad) cudaFree(h->dec_bad);
  for (cudaEvent_t e : h->prof_pool) cudaEventDestroy(e);
  for (auto& kv : h->taps) { cudaFree(kv.second.d); if (kv.second.g) cudaFree(kv.second.g); }
  delete h;
}

int mimi_b200_debug_set(mimi_b200_t* h, int key, int value) {
  if (!h) return MIMI_B200_ERR_ARG;
  if (key == 0) h->dbg_layers = std::min(std::max(value, 0), MIMI_B200_NUM_LAYERS);
  else if (key == 1) h->dbg_last_conv = std::min(std::max(value, 0), MIMI_B200_NUM_CONVS - 1);
  else if (key == 2) { h->prof_on = value != 0; h->prof_n = 0; }
  else if (key == 3) {
    if (value != 0 && value != 7 && value != 9) return fail(h, MIMI_B200_ERR_ARG, "debug_set: kernel generation must be 0, 7 or 9");
    h->mode = value;
  }
  else if (key == 5) h->exp_chunk_kb = std::max(value, 0);
  else if (key == 9) { h->exp_pair_n128 = std::max(value, 0); h->amap_cache.clear(); }
  else if (key == 10) h->exp_no_flat = value != 0;
  else if (key == 11) h->exp_linear_k = value != 0;      // 1: k-blocks in linear order (no tap grouping)
  else if (key == 13) h->exp_no_tile_list = value != 0;
  else if (key == 16) h->exp_resample_simple = value != 0;
  else if (key == 17) h->exp_front_tf32 = value != 0;
  else if (key == 18) h->exp_att_grid = value != 0;
  else if (key == 19) h->exp_rvq_tf32 = value != 0;
  else if (key == 21) h->exp_gelu_erff = value != 0;
  else if (key == 20) { h->exp_no_taps = value != 0; h->amap_cache.clear(); }     // (the activation maps of a conv differ)
  else return fail(h, MIMI_B200_ERR_ARG, "debug_set: unknown key");
  return MIMI_B200_OK;
}

int mimi_b200_profile_read(mimi_b200_t* h, int max_ids, double* sum_ms, int64_t* count) {
  if (!h || !sum_ms || !count || max_ids <= 0) return fail(h, MIMI_B200_ERR_ARG, "profile_read: bad argument");
  for (int i = 0; i < max_ids; ++i) { sum_ms[i] = 0.0; count[i] = 0; }
  if (h->prof_n == 0) return MIMI_B200_OK;
  CUDA_TRY(h, cudaEventSynchronize(h->prof_pool[h->prof_n - 1]));
  for (size_t i = 1; i < h->prof_n; ++i) {
    const int id = h->prof_ids[i];
    if (id < 0 || id >= max_ids) continue;            // -1 = begin marker of an encode call
    float ms = 0.f;
    CUDA_TRY(h, cudaEventElapsedTime(&ms, h->prof_pool[i - 1], h->prof_pool[i]));
    sum_ms[id] += ms;
    count[id] += 1;
  }
  h->prof_n = 0;
  return MIMI_B200_OK;
}

int mimi_b200_load_weights(mimi_b200_t* h, const mimi_b200_weights_t* w) {
  if (!h || !w) return fail(h, MIMI_B200_ERR_ARG, "load_weights: NULL argument");
  DeviceGuard guard;
  CUDA_TRY(h, cudaSetDevice(h->device));
  free_weights(h);
  for (int i = 0; i < MIMI_B200_NUM_CONVS; ++i)
    if (!w->conv_weight[i] || !w->conv_bias[i]) return fail(h, MIMI_B200_ERR_ARG, "load_weights: missing SEANet conv tensor");
  int rc;
  for (int i = 0; i < MIMI_B200_NUM_CONVS; ++i) {
    const ConvGeom& g = kConv[i];
    std::vector<float> wt = (i == 0) ? std::vector<float>(w->conv_weight[0], w->conv_weight[0] + 64 * 7)
                                     : pack_conv(w->conv_weight[i], g.cout, g.cin, g.k);
    if ((rc = dev_upload(h, &h->conv_wt[i], wt))) return rc;
    if ((rc = dev_upload(h, &h->conv_b[i], std::vector<float>(w->conv_bias[i], w->conv_bias[i] + g.cout)))) return rc;
  }
  for (int l = 0; l < MIMI_B200_NUM_LAYERS; ++l) {
    const mimi_b200_layer_weights_t& s = w->layer[l];
    const float* need[] = {s.input_layernorm_weight, s.input_layernorm_bias, s.q_proj_weight, s.k_proj_weight,
                           s.v_proj_weight, s.o_proj_weight, s.self_attn_layer_scale, s.post_attention_layernorm_weight,
                           s.post_attention_layernorm_bias, s.fc1_weight, s.fc2_weight, s.mlp_layer_scale};
    for (const float* q : need) if (!q) return fail(h, MIMI_B200_ERR_ARG, "load_weights: missing transformer tensor");
    LayerDev& d = h->layer[l];
    auto vec = [](const float* p, size_t n) { return std::vector<float>(p, p + n); };
    if ((rc = dev_upload(h, &d.ln1_w, vec(s.input_layernorm_weight, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ln1_b, vec(s.input_layernorm_bias, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ln2_w, vec(s.post_attention_layernorm_weight, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ln2_b, vec(s.post_attention_layernorm_bias, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ls1, vec(s.self_attn_layer_scale, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ls2, vec(s.mlp_layer_scale, 512)))) return rc;
    // fused QKV: Wt [512][1536], columns [q | k | v]
    std::vector<float> qkv((size_t)512 * 1536);
    const float* src[3] = {s.q_proj_weight, s.k_proj_weight, s.v_proj_weight};
    for (int part = 0; part < 3; ++part)
      for (int n = 0; n < 512; ++n)
        for (int k = 0; k < 512; ++k) qkv[(size_t)k * 1536 + part * 512 + n] = src[part][(size_t)n * 512 + k];
    if ((rc = dev_upload(h, &d.qkv_wt, qkv))) return rc;
    if ((rc = dev_upload(h, &d.o_wt, transpose_nk(s.o_proj_weight, 512, 512)))) return rc;
    if ((rc = dev_upload(h, &d.fc1_wt, transpose_nk(s.fc1_weight, 2048, 512)))) return rc;
    if ((rc = dev_upload(h, &d.fc2_wt, transpose_nk(s.fc2_weight, 512, 2048)))) return rc;
  }
  if (!w->downsample_weight || !w->semantic_input_proj_weight || !w->acoustic_input_proj_weight)
    return fail(h, MIMI_B200_ERR_ARG, "load_weights: missing downsample / input_proj tensor");
  if ((rc = dev_upload(h, &h->down_wt, pack_conv(w->downsample_weight, 512, 512, 4)))) return rc;
  {
    std::vector<float> pj((size_t)512 * 512);
    for (int n = 0; n < 256; ++n)
      for (int k = 0; k < 512; ++k) {
        pj[(size_t)k * 512 + n] = w->semantic_input_proj_weight[(size_t)n * 512 + k];
        pj[(size_t)k * 512 + 256 + n] = w->acoustic_input_proj_weight[(size_t)n * 512 + k];
      }
    if ((rc = dev_upload(h, &h->proj_wt, pj))) return rc;
  }
  {
    // MimiEuclideanCodebook.embed (modeling_mimi.py:1191-1195): embed_sum / clamp(cluster_usage, 1e-5)
    const size_t per = (size_t)kCodebookSize * kCodeDim;
    std::vector<float> E(32 * per), Et(32 * per), En((size_t)32 * kCodebookSize);
    for (int s = 0; s < MIMI_B200_MAX_QUANTIZERS; ++s) {
      if (!w->embed_sum[s] || !w->cluster_usage[s]) return fail(h, MIMI_B200_ERR_ARG, "load_weights: missing codebook tensor");
      for (int c = 0; c < kCodebookSize; ++c) {
        const float u = std::max(w->cluster_usage[s][c], 1e-5f);
        double nrm = 0.0;
        for (int k = 0; k < kCodeDim; ++k) {
          const float v = w->embed_sum[s][(size_t)c * kCodeDim + k] / u;
          E[s * per + (size_t)c * kCodeDim + k] = v;
          Et[s * per + (size_t)k * kCodebookSize + c] = v;
          nrm += (double)v * (double)v;
        }
        En[(size_t)s * kCodebookSize + c] = (float)nrm;
      }
    }
    if ((rc = dev_upload(h, &h->embed, E))) return rc;
    if ((rc = dev_upload(h, &h->embed_t, Et))) return rc;
    if ((rc = dev_upload(h, &h->enorm, En))) return rc;
    {
      // fp16 generation: rows scaled by 2^s so that max |e'| lies in [2^13, 2^14), e' = hi16 + lo16; the distance's -2 carries 2^-s
      const size_t tot = 32 * per;
      std::vector<uint16_t> f(2 * tot);
      std::vector<float> m2s((size_t)32 * kCodebookSize);
      auto bits = [](float v) { const __half hh = __float2half_rn(v); uint16_t u; memcpy(&u, &hh, 2); return u; };
      for (size_t r = 0; r < (size_t)32 * kCodebookSize; ++r) {
        float mx = 0.f;
        for (int k = 0; k < kCodeDim; ++k) mx = std::max(mx, std::fabs(E[r * kCodeDim + k]));
        int e = 0;
        if (mx > 0.f && std::isfinite(mx)) { int ex; std::frexp(mx, &ex); e = 14 - ex; }
        m2s[r] = -2.0f * std::ldexp(1.0f, -e);
        for (int k = 0; k < kCodeDim; ++k) {
          const float v = std::ldexp(E[r * kCodeDim + k], e);
          const float vh = __half2float(__float2half_rn(v));
          f[r * kCodeDim + k] = bits(vh);
          f[tot + r * kCodeDim + k] = bits(v - vh);
        }
      }
      CUDA_TRY(h, cudaMalloc((void**)&h->embed16, f.size() * sizeof(uint16_t)));
      h->allocs.push_back(h->embed16);
      CUDA_TRY(h, cudaMemcpy(h->embed16, f.data(), f.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
      if ((rc = dev_upload(h, &h->embed_m2s, m2s))) return rc;
    }
    for (size_t i = 0; i < E.size(); ++i) { float hi, lo; split_tf32(E[i], hi, lo); E[i] = hi; Et[i] = lo; }   // reuse as hi / lo
    if ((rc = dev_upload(h, &h->embed_hi, E))) return rc;
    if ((rc = dev_upload(h, &h->embed_lo, Et))) return rc;
  }
  {
    // MimiRotaryEmbedding (modeling_mimi.py:538-577): angle = float(pos) * inv_freq[i] in fp32
    std::vector<float> cs((size_t)kRopeMaxPos * 32), sn((size_t)kRopeMaxPos * 32);
    float inv[32];
    for (int i = 0; i < 32; ++i)
      inv[i] = w->rope_inv_freq ? w->rope_inv_freq[i] : 1.0f / powf(10000.0f, (float)(2 * i) / 64.0f);
    for (int pos = 0; pos < kRopeMaxPos; ++pos)
      for (int i = 0; i < 32; ++i) {
        const float ang = (float)pos * inv[i];
        cs[(size_t)pos * 32 + i] = cosf(ang);
        sn[(size_t)pos * 32 + i] = sinf(ang);
      }
    if ((rc = dev_upload(h, &h->rope_cos, cs))) return rc;
    if ((rc = dev_upload(h, &h->rope_sin, sn))) return rc;
  }
  std::memcpy(h->f0_consts.w0, w->conv_weight[0], sizeof(float) * 64 * 7);
  std::memcpy(h->f0_consts.b0, w->conv_bias[0], sizeof(float) * 64);
  std::memcpy(h->f0_consts.b1, w->conv_bias[1], sizeof(float) * 32);
  std::memcpy(h->f0_consts.b2, w->conv_bias[2], sizeof(float) * 64);
  if ((rc = tc_load_weights(h, w))) return rc;
  {
    const cuuint64_t dims[2] = {(cuuint64_t)kCodeDim, (cuuint64_t)MIMI_B200_MAX_QUANTIZERS * kCodebookSize};
    const cuuint64_t strides[1] = {(cuuint64_t)kCodeDim * sizeof(float)};
    if ((rc = tc_make_map(h, &h->map_embed_hi, h->embed_hi, 2, dims, strides, rvqtc::kCodesPerBlock))) return rc;
    if ((rc = tc_make_map(h, &h->map_embed_lo, h->embed_lo, 2, dims, strides, rvqtc::kCodesPerBlock))) return rc;
    const cuuint64_t strides16[1] = {(cuuint64_t)kCodeDim * sizeof(uint16_t)};
    const size_t tot = (size_t)MIMI_B200_MAX_QUANTIZERS * kCodebookSize * kCodeDim;
    if ((rc = tc_make_map16_sw128(h, &h->map_embed16_hi, h->embed16, dims, strides16, rvq16::kCodesPerBlock))) return rc;
    if ((rc = tc_make_map16_sw128(h, &h->map_embed16_lo, h->embed16 + tot, dims, strides16, rvq16::kCodesPerBlock))) return rc;
  }
  h->amap_cache.clear();
  h->loaded = true;
  return MIMI_B200_OK;
}

int mimi_b200_workspace_bytes(mimi_b200_t* h, int B, int64_t N, int K, size_t* out_bytes) {
  if (!h || !out_bytes) return fail(h, MIMI_B200_ERR_ARG, "workspace_bytes: NULL argument");
  if (B < 0 || N < 0 || N > (1ll << 30) || K < 1 || K > MIMI_B200_MAX_QUANTIZERS)
    return fail(h, MIMI_B200_ERR_ARG, "workspace_bytes: bad B/N/K");
  // sized for the kernel generation in force (debug_set key 3)
  const bool simt = h->mode == 0 || h->dbg_last_conv != MIMI_B200_NUM_CONVS - 1;
  if (simt) { *out_bytes = make_plan(B, N, K).bytes + 256; return MIMI_B200_OK; }
  const PlanTC pt = make_plan_tc(B, N, K, h->mode == 9);
  *out_bytes = pt.bytes + 256;
  return ensure_stage(h, (pt.bytes - (size_t)pt.ints) / sizeof(int));     // lengths + tile lists of a batch this size
}

static int encode_impl(mimi_b200_t* h, const float* d_input, int B, int64_t N, const int64_t* h_valid_len, int K,
                       int64_t* d_codes, float* d_latent_opt, void* d_workspace, size_t workspace_bytes, void* stream);

int mimi_b200_encode(mimi_b200_t* h, const float* d_input, int B, int64_t N, const int64_t* h_valid_len, int K,
                     int64_t* d_codes, float* d_latent_opt, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (!h) return MIMI_B200_ERR_ARG;
  h->phase = 0; h->front_b0 = 0; h->front_b1 = B;
  return encode_impl(h, d_input, B, N, h_valid_len, K, d_codes, d_latent_opt, d_workspace, workspace_bytes, stream);
}

int mimi_b200_encode_phase(mimi_b200_t* h, int phase, int b0, int b1, const float* d_input, int B, int64_t N,
                           const int64_t* h_valid_len, int K, int64_t* d_codes, float* d_latent_opt, void* d_workspace,
                           size_t workspace_bytes, void* stream) {
  if (!h) return MIMI_B200_ERR_ARG;
  if (phase < MIMI_B200_PHASE_BEGIN || phase > MIMI_B200_PHASE_FINISH) return fail(h, MIMI_B200_ERR_ARG, "encode_phase: bad phase");
  if (h->mode == 0 || h->dbg_last_conv != MIMI_B200_NUM_CONVS - 1)
    return fail(h, MIMI_B200_ERR_STATE, "encode_phase: needs a tensor-core generation (mode 7 or 9)");
  if (phase == MIMI_B200_PHASE_FRONT && (b0 < 0 || b1 > B || b0 > b1)) return fail(h, MIMI_B200_ERR_ARG, "encode_phase: bad item range");
  h->phase = phase; h->front_b0 = b0; h->front_b1 = b1;
  const int rc = encode_impl(h, d_input, B, N, h_valid_len, K, d_codes, d_latent_opt, d_workspace, workspace_bytes, stream);
  h->phase = 0;
  return rc;
}

static int encode_impl(mimi_b200_t* h, const float* d_input, int B, int64_t N, const int64_t* h_valid_len, int K,
                       int64_t* d_codes, float* d_latent_opt, void* d_workspace, size_t workspace_bytes, void* stream) {
  if (!h) return MIMI_B200_ERR_ARG;
  if (!h->loaded) return fail(h, MIMI_B200_ERR_STATE, "encode: weights not loaded");
  if (K > MIMI_B200_MAX_QUANTIZERS)
    return fail(h, MIMI_B200_ERR_ARG,
                "The number of quantizers (i.e codebooks) asked should be lower than the total number of quantizers 32, "
                "but is currently " + std::to_string(K) + ".");
  if (K < 1)
    return fail(h, MIMI_B200_ERR_ARG,
                "The number of quantizers (i.e codebooks) asked should be higher than the number of semantic quantizers 1, "
                "but is currently " + std::to_string(K) + ".");
  if (B < 0 || N < 0 || N > (1ll << 30)) return fail(h, MIMI_B200_ERR_ARG, "encode: bad B/N");
  if (B == 0 || N == 0) return MIMI_B200_OK;
  if (!d_input || !d_codes || !d_workspace) return fail(h, MIMI_B200_ERR_ARG, "encode: NULL device pointer");
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  const bool use_tc = h->mode >= 1 && h->dbg_last_conv == MIMI_B200_NUM_CONVS - 1;
  const Plan p = make_plan(B, N, K);
  const PlanTC pt = make_plan_tc(B, N, K, h->mode == 9);
  const size_t need = use_tc ? pt.bytes : p.bytes;
  // align the workspace base to 256 bytes
  uintptr_t base = (reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255);
  if (base + need > reinterpret_cast<uintptr_t>(d_workspace) + workspace_bytes)
    return fail(h, MIMI_B200_ERR_WORKSPACE, "encode: workspace too small, need " + std::to_string(need + 256));
  if (p.rows[4] > kRopeMaxPos) return fail(h, MIMI_B200_ERR_ARG, "encode: more than 65536 25-Hz positions (43.7 min) per item");
  float* ws = reinterpret_cast<float*>(base);
  int* dints = reinterpret_cast<int*>(base + (use_tc ? pt.ints : p.ints));
  h->last = p;
  h->last_tc = pt;
  h->last_was_tc = use_tc;
  h->last_mode = h->mode;
  h->last_ws = ws;
  mark(h, -1, st);

  // per-level lengths. strict: uniform (no arrays). ragged: item i is encoded over
  // min(N, ceil(len_i/1920)*1920) samples (identical kept frames, see mimi_b200.h).
  const int* dlen[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  const int* dprefix = nullptr;
  int maxlen[6];
  for (int l = 0; l < 6; ++l) maxlen[l] = p.rows[l];
  int total_frames = B * p.rows[5];
  for (int l = 0; l < 6; ++l) h->item_tiles[l] = (long long)B * ((p.rows[l] + 127) / 128);
  for (int l = 0; l < 6; ++l) { h->tile_ptr[l] = nullptr; h->tile_cnt[l] = 0; }
  h->len0_host.assign((size_t)B, (int)N);
  if (h_valid_len) {
    for (int l = 0; l < 6; ++l) h->item_tiles[l] = 0;
    std::vector<int> v((size_t)7 * B + 1);
    v.reserve((pt.bytes - (size_t)pt.ints) / sizeof(int));
    int mx[6] = {0, 0, 0, 0, 0, 0};
    int acc = 0;
    for (int b = 0; b < B; ++b) {
      long long len = h_valid_len[b];
      if (len < 0 || len > N) return fail(h, MIMI_B200_ERR_ARG, "encode: valid_len out of range");
      long long L = std::min<long long>(N, (len + MIMI_B200_FRAME_SIZE - 1) / MIMI_B200_FRAME_SIZE * MIMI_B200_FRAME_SIZE);
      h->len0_host[b] = (int)L;
      for (int l = 0; l < 6; ++l) {
        v[(size_t)l * B + b] = (int)L;
        mx[l] = std::max(mx[l], (int)L);
        h->item_tiles[l] += (L + 127) / 128;
        if (l < 5) L = (L + kLevelStride[l] - 1) / kLevelStride[l];
      }
      v[(size_t)6 * B + b] = acc;
      acc += v[(size_t)5 * B + b];
    }
    v[(size_t)7 * B] = acc;
    total_frames = acc;
    for (int l = 0; l < 6; ++l) { h->tile_ptr[l] = nullptr; h->tile_cnt[l] = 0; }
    if (use_tc && !h->exp_no_tile_list && B < 2048) {
      // compact lists of the 128-row tiles that exist at levels 1..5 (the outputs of every conv), m-tile major so that
      // neighbouring CTAs work on the same stretch of time; the persistent GEMMs deal them out round-robin
      for (int l = 1; l < 6; ++l) {
        const size_t off = v.size();
        const int mt_max = (mx[l] + 127) / 128;
        for (int mt = 0; mt < mt_max; ++mt)
          for (int b = 0; b < B; ++b)
            if (mt * 128 < v[(size_t)l * B + b]) v.push_back((b << 20) | mt);
        h->tile_ptr[l] = dints + off;
        h->tile_cnt[l] = (int)(v.size() - off);
      }
    }
    if (h->phase <= MIMI_B200_PHASE_BEGIN) {       // the later phases of a phased call find the lengths in the workspace
      int rc = stage_ints(h, v, dints, st);
      if (rc) return rc;
    }
    for (int l = 0; l < 6; ++l) { dlen[l] = dints + (size_t)l * B; maxlen[l] = mx[l]; }
    dprefix = dints + (size_t)6 * B;
    if (h->phase <= MIMI_B200_PHASE_BEGIN) {
      const long long ncodes = (long long)B * K * p.rows[5];
      fill_codes_zero_kernel<<<(unsigned)((ncodes + 255) / 256), 256, 0, st>>>(reinterpret_cast<long long*>(d_codes), ncodes);
      h->launches++; mark(h, 24, st);
    }
  }

  if (use_tc) {
    int rc_tc = encode_tc(h, d_input, B, N, K, pt, ws, dlen, maxlen, dprefix, total_frames, d_codes, d_latent_opt, st);
    if (rc_tc == MIMI_B200_OK && !h->sync_err.empty()) return fail(h, MIMI_B200_ERR_CUDA, h->sync_err);
    return rc_tc;
  }

  auto istride = [&](int level, int C) { return (long long)p.rows[level] * C; };
  int rc;

  // ---- SEANet encoder (MimiEncoder.forward, modeling_mimi.py:490-496) ------------------------------
  {
    dim3 grid((maxlen[0] + 127) / 128, B);
    if (maxlen[0] > 0) {
      conv0_kernel<<<grid, 256, 0, st>>>(d_input, N, h->conv_wt[0], h->conv_b[0], ws + p.a0, istride(0, 64), dlen[0], maxlen[0]);
      h->launches++; mark(h, 0, st);
      CUDA_TRY(h, cudaGetLastError());
    }
  }
  struct Stage { long long h_off, r_off; int level, C; };
  const Stage stages[4] = {{p.a0, p.r1, 0, 64}, {p.d1, p.r2, 1, 128}, {p.d2, p.r3, 2, 256}, {p.d3, p.r4, 3, 512}};
  const long long down_off[4] = {p.d1, p.d2, p.d3, p.d4};
  bool stop = h->dbg_last_conv < 1;
  for (int s = 0; s < 4 && !stop; ++s) {
    const Stage& sg = stages[s];
    const int ia = 1 + 3 * s, ib = 2 + 3 * s, id = 3 + 3 * s;
    GemmParams g{};
    // resblock conv a: ELU -> C -> C/2, k3
    g.A = ws + sg.h_off; g.Wt = h->conv_wt[ia]; g.bias = h->conv_b[ia]; g.out = ws + sg.r_off;
    g.len_in = dlen[sg.level]; g.uniform_len_in = maxlen[sg.level];
    g.a_item_stride = istride(sg.level, sg.C); g.out_item_stride = istride(sg.level, sg.C / 2);
    g.Cin = sg.C; g.stride = 1; g.pad_left = 2; g.K = 3 * sg.C; g.N = sg.C / 2; g.elu_in = 1;
    if ((rc = launch_gemm(h, g, B, maxlen[sg.level], st, ia))) return rc;
    if (h->dbg_last_conv <= ia) { stop = true; break; }
    // resblock conv b: ELU -> C/2 -> C, k1, + skip (in place on h)
    g = GemmParams{};
    g.A = ws + sg.r_off; g.Wt = h->conv_wt[ib]; g.bias = h->conv_b[ib]; g.res = ws + sg.h_off; g.out = ws + sg.h_off;
    g.len_in = dlen[sg.level]; g.uniform_len_in = maxlen[sg.level];
    g.a_item_stride = istride(sg.level, sg.C / 2); g.out_item_stride = istride(sg.level, sg.C);
    g.Cin = sg.C / 2; g.stride = 1; g.pad_left = 0; g.K = sg.C / 2; g.N = sg.C; g.elu_in = 1;
    if ((rc = launch_gemm(h, g, B, maxlen[sg.level], st, ib))) return rc;
    if (h->dbg_last_conv <= ib) { stop = true; break; }
    // strided down conv: ELU -> C -> 2C, k = 2*ratio, stride = ratio
    const ConvGeom& cg = kConv[id];
    g = GemmParams{};
    g.A = ws + sg.h_off; g.Wt = h->conv_wt[id]; g.bias = h->conv_b[id]; g.out = ws + down_off[s];
    g.len_in = dlen[sg.level]; g.uniform_len_in = maxlen[sg.level];
    g.a_item_stride = istride(sg.level, sg.C); g.out_item_stride = istride(sg.level + 1, 2 * sg.C);
    g.Cin = sg.C; g.stride = cg.stride; g.pad_left = cg.k - cg.stride; g.K = cg.k * sg.C; g.N = 2 * sg.C; g.elu_in = 1;
    if ((rc = launch_gemm(h, g, B, maxlen[sg.level], st, id))) return rc;
    if (h->dbg_last_conv <= id) { stop = true; break; }
  }
  if (stop) return MIMI_B200_OK;
  {
    GemmParams g{};   // final conv: ELU -> 1024 -> 512, k3
    g.A = ws + p.d4; g.Wt = h->conv_wt[13]; g.bias = h->conv_b[13]; g.out = ws + p.z;
    g.len_in = dlen[4]; g.uniform_len_in = maxlen[4];
    g.a_item_stride = istride(4, 1024); g.out_item_stride = istride(4, 512);
    g.Cin = 1024; g.stride = 1; g.pad_left = 2; g.K = 3072; g.N = 512; g.elu_in = 1;
    if ((rc = launch_gemm(h, g, B, maxlen[4], st, 13))) return rc;
  }

  // ---- encoder transformer (MimiTransformerModel.forward, modeling_mimi.py:1015-1140) ---------------
  const int T25 = maxlen[4];
  for (int l = 0; l < h->dbg_layers; ++l) {
    const LayerDev& d = h->layer[l];
    dim3 lgrid((T25 + 7) / 8, B);
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z, ws + p.y, d.ln1_w, d.ln1_b, istride(4, 512), dlen[4], T25, nullptr);
    h->launches++; mark(h, 14, st);
    GemmParams g{};
    g.A = ws + p.y; g.Wt = d.qkv_wt; g.out = ws + p.qkv; g.len_in = dlen[4]; g.uniform_len_in = T25;
    g.a_item_stride = istride(4, 512); g.out_item_stride = istride(4, 1536);
    g.Cin = 512; g.stride = 1; g.K = 512; g.N = 1536;
    if ((rc = launch_gemm(h, g, B, T25, st, 15))) return rc;
    dim3 agrid((T25 + kAttQT - 1) / kAttQT, kHeads, B);
    swa_attention_kernel<<<agrid, 256, kAttSmemBytes, st>>>(ws + p.qkv, istride(4, 1536), ws + p.att, istride(4, 512),
                                                            h->rope_cos, h->rope_sin, dlen[4], T25);
    h->launches++; mark(h, 16, st);
    CUDA_TRY(h, cudaGetLastError());
    g = GemmParams{};   // o_proj + LayerScale + residual (in place on z)
    g.A = ws + p.att; g.Wt = d.o_wt; g.scale = d.ls1; g.res = ws + p.z; g.out = ws + p.z;
    g.len_in = dlen[4]; g.uniform_len_in = T25; g.a_item_stride = istride(4, 512); g.out_item_stride = istride(4, 512);
    g.Cin = 512; g.stride = 1; g.K = 512; g.N = 512;
    if ((rc = launch_gemm(h, g, B, T25, st, 17))) return rc;
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z, ws + p.y, d.ln2_w, d.ln2_b, istride(4, 512), dlen[4], T25, nullptr);
    h->launches++; mark(h, 14, st);
    g = GemmParams{};   // fc1 + GELU(erf)
    g.A = ws + p.y; g.Wt = d.fc1_wt; g.out = ws + p.ffn; g.len_in = dlen[4]; g.uniform_len_in = T25;
    g.a_item_stride = istride(4, 512); g.out_item_stride = istride(4, 2048);
    g.Cin = 512; g.stride = 1; g.K = 512; g.N = 2048; g.act = 1;
    if ((rc = launch_gemm(h, g, B, T25, st, 18))) return rc;
    g = GemmParams{};   // fc2 + LayerScale + residual (in place on z)
    g.A = ws + p.ffn; g.Wt = d.fc2_wt; g.scale = d.ls2; g.res = ws + p.z; g.out = ws + p.z;
    g.len_in = dlen[4]; g.uniform_len_in = T25; g.a_item_stride = istride(4, 2048); g.out_item_stride = istride(4, 512);
    g.Cin = 2048; g.stride = 1; g.K = 2048; g.N = 512;
    if ((rc = launch_gemm(h, g, B, T25, st, 19))) return rc;
  }

  // ---- stride-2 downsample conv, replicate padding (modeling_mimi.py:1422-1431,1484) ----------------
  {
    GemmParams g{};
    g.A = ws + p.z; g.Wt = h->down_wt; g.out = ws + p.e; g.len_in = dlen[4]; g.uniform_len_in = T25;
    g.a_item_stride = istride(4, 512); g.out_item_stride = istride(5, 512);
    g.Cin = 512; g.stride = 2; g.pad_left = 2; g.K = 2048; g.N = 512; g.replicate = 1;
    if ((rc = launch_gemm(h, g, B, T25, st, 20))) return rc;
  }
  const int T = maxlen[5];
  if (d_latent_opt) {
    const long long n = (long long)kHidden * p.rows[5];
    dim3 tgrid((unsigned)((n + 255) / 256), B);
    latent_transpose_kernel<<<tgrid, 256, 0, st>>>(ws + p.e, istride(5, 512), d_latent_opt, p.rows[5], dlen[5], T);
    h->launches++; mark(h, 23, st);
  }
  // ---- split RVQ (modeling_mimi.py:1311-1338): both input_proj as one GEMM, then the fused chain -----
  {
    GemmParams g{};
    g.A = ws + p.e; g.Wt = h->proj_wt; g.out = ws + p.rp; g.len_in = dlen[5]; g.uniform_len_in = T;
    g.a_item_stride = istride(5, 512); g.out_item_stride = istride(5, 512);
    g.Cin = 512; g.stride = 1; g.K = 512; g.N = 512;
    if ((rc = launch_gemm(h, g, B, T, st, 21))) return rc;
    RvqParams r{};
    r.rproj = ws + p.rp; r.item_stride = istride(5, 512);
    r.embed = h->embed; r.embed_t = h->embed_t; r.enorm = h->enorm;
    r.codes = reinterpret_cast<long long*>(d_codes); r.K = K; r.T_out = p.rows[5];
    r.len = dlen[5]; r.uniform_len = T; r.B = B; r.total_frames = total_frames; r.frame_prefix = dprefix;
    if (total_frames > 0) {
      rvq_encode_kernel<<<(total_frames + kRvqFM - 1) / kRvqFM, 256, kRvqSmemBytes, st>>>(r);
      h->launches++; mark(h, 22, st);
    }
  }
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}

int mimi_b200_debug_tap(mimi_b200_t* h, int which, float* d_out, size_t cap, int64_t* rows_per_item, int* channels,
                        void* stream) {
  if (!h || !h->last_ws) return fail(h, MIMI_B200_ERR_STATE, "debug_tap: no encode call yet");
  const Plan& p = h->last;
  TapInfo t{};
  if (h->last_was_tc) {
    const PlanTC& q = h->last_tc;
    if (which < 3) return fail(h, MIMI_B200_ERR_ARG, "debug_tap: level-0 activations stay on chip in the tensor-core generations");
    switch (which) {
      case 3: t = {q.d1, 1, 128}; break;
      case 6: t = {q.d2, 2, 256}; break;
      case 9: t = {q.d3, 3, 512}; break;
      case 13: t = {q.z, 4, 512}; break;
      case 200: t = {q.e, 5, 512}; break;
      case 201: t = {q.rp, 5, 512}; break;
      default:
        if (which >= 100 && which < 100 + MIMI_B200_NUM_LAYERS) t = {q.z, 4, 512};
        else return fail(h, MIMI_B200_ERR_ARG, "debug_tap: tap not materialised on the tensor-core path");
    }
  } else
  switch (which) {
    case 0: case 2: t = {p.a0, 0, 64}; break;
    case 1: t = {p.r1, 0, 32}; break;
    case 3: case 5: t = {p.d1, 1, 128}; break;
    case 4: t = {p.r2, 1, 64}; break;
    case 6: case 8: t = {p.d2, 2, 256}; break;
    case 7: t = {p.r3, 2, 128}; break;
    case 9: case 11: t = {p.d3, 3, 512}; break;
    case 10: t = {p.r4, 3, 256}; break;
    case 12: t = {p.d4, 4, 1024}; break;
    case 13: t = {p.z, 4, 512}; break;
    case 200: t = {p.e, 5, 512}; break;
    case 201: t = {p.rp, 5, 512}; break;
    default:
      if (which >= 100 && which < 100 + MIMI_B200_NUM_LAYERS) t = {p.z, 4, 512};
      else return fail(h, MIMI_B200_ERR_ARG, "debug_tap: unknown tap");
  }
  const size_t n = (size_t)p.B * p.rows[t.level] * t.C;
  if (rows_per_item) *rows_per_item = p.rows[t.level];
  if (channels) *channels = t.C;
  if (!d_out) return MIMI_B200_OK;
  if (cap < n) return fail(h, MIMI_B200_ERR_ARG, "debug_tap: output too small");
  CUDA_TRY(h, cudaMemcpyAsync(d_out, static_cast<float*>(h->last_ws) + t.off, n * sizeof(float), cudaMemcpyDeviceToDevice,
                              static_cast<cudaStream_t>(stream)));
  return MIMI_B200_OK;
}

// Unit-test hook for the tensor-core GEMM: out[M][N] = epilogue(A[M][K] * W[N][K]^T) through the same split /
// TMA / tcgen05 path the encoder uses (A is split on the device, W on the host), in the generation selected by key 3.
__global__ void debug_split_kernel(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, long long n, int lob) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i < n) store_split4_x(hi, lo, i, *reinterpret_cast<const float4*>(x + i), lob);
}

int mimi_b200_debug_tc_gemm(mimi_b200_t* h, const float* d_a, const float* h_w, const float* d_bias_opt, int M, int N,
                            int K, int act, float* d_out, void* stream) {
  if (!h || !d_a || !h_w || !d_out) return fail(h, MIMI_B200_ERR_ARG, "debug_tc_gemm: NULL argument");
  if (h->mode != 7 && h->mode != 9) return fail(h, MIMI_B200_ERR_STATE, "debug_tc_gemm: needs a tensor-core generation (mode 7 or 9)");
  if (M <= 0 || N % 64 || K % 32) return fail(h, MIMI_B200_ERR_ARG, "debug_tc_gemm: need N % 64 == 0 and K % 32 == 0");
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = tc_init_driver(h))) return rc;
  const size_t mark = h->allocs.size();          // the test weight is freed again below
  TcWeight w;
  std::vector<float> hbias;
  if (d_bias_opt) {
    hbias.resize(N);
    CUDA_TRY(h, cudaMemcpy(hbias.data(), d_bias_opt, (size_t)N * sizeof(float), cudaMemcpyDeviceToHost));
  }
  if ((rc = tc_make_weight(h, &w, std::vector<float>(h_w, h_w + (size_t)N * K), N, K, d_bias_opt ? hbias.data() : nullptr))) return rc;
  const bool f16 = h->mode == 9;
  float *hi = nullptr, *lo = nullptr;
  const long long n = (long long)M * K;
  CUDA_TRY(h, cudaMalloc((void**)&hi, n * sizeof(float)));
  CUDA_TRY(h, cudaMalloc((void**)&lo, n * sizeof(float)));
  debug_split_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(d_a, hi, lo, n, f16 ? 3 : 1);
  CUtensorMap ma_hi, ma_lo;
  const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)M, 1};
  const cuuint64_t strides[2] = {(cuuint64_t)K * sizeof(float), (cuuint64_t)n * sizeof(float)};
  const cuuint64_t strides_b[2] = {(cuuint64_t)K * 2, (cuuint64_t)n * 2};
  if (f16) { if ((rc = tc_make_map_bf16(h, &ma_hi, hi, 3, dims, strides_b, tc::kBM))) return rc; }
  else if ((rc = tc_make_map(h, &ma_hi, hi, 3, dims, strides, tc::kBM))) return rc;
  if ((rc = tc_make_map_bf16(h, &ma_lo, lo, 3, dims, strides_b, tc::kBM))) return rc;
  tc::Epilogue ep{};
  ep.cmul = w.cmul[f16 ? 1 : 0]; ep.cadd = w.cadd; ep.out_raw = d_out; ep.raw_item_stride = (long long)M * N; ep.act = act;
  ep.uniform_len_in = M; ep.conv_stride = 1; ep.N = N;
  ep.chunk_kb = h->exp_chunk_kb; ep.lo_bf16 = f16 ? 3 : 1;
  rc = launch_tcp(h, ma_hi, ma_lo, w, ep, 1, (M + tc::kBM - 1) / tc::kBM, st);
  h->launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(hi);
  cudaFree(lo);
  for (size_t i = mark; i < h->allocs.size(); ++i) cudaFree(h->allocs[i]);
  h->allocs.resize(mark);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(h, MIMI_B200_ERR_CUDA, std::string("debug_tc_gemm: ") + cudaGetErrorString(e));
  return MIMI_B200_OK;
}

int mimi_b200_debug_shift_probe(mimi_b200_t* h, const float* d_a, const float* h_w, int K, int shift, int base_mode,
                                float* d_out, void* stream) {
  if (!h || !d_a || !h_w || !d_out) return fail(h, MIMI_B200_ERR_ARG, "debug_shift_probe: NULL argument");
  if (K <= 0 || K % 32 || shift < 0 || shift > 8) return fail(h, MIMI_B200_ERR_ARG, "debug_shift_probe: bad K / shift");
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  if ((rc = tc_init_driver(h))) return rc;
  float* dw = nullptr;
  CUDA_TRY(h, cudaMalloc((void**)&dw, (size_t)64 * K * sizeof(float)));
  CUDA_TRY(h, cudaMemcpy(dw, h_w, (size_t)64 * K * sizeof(float), cudaMemcpyHostToDevice));
  CUtensorMap ma, mw;
  const cuuint64_t adims[2] = {(cuuint64_t)K, 136};
  const cuuint64_t wdims[2] = {(cuuint64_t)K, 64};
  const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  if ((rc = tc_make_map(h, &ma, d_a, 2, adims, strides, 136))) { cudaFree(dw); return rc; }
  if ((rc = tc_make_map(h, &mw, dw, 2, wdims, strides, 64))) { cudaFree(dw); return rc; }
  tc2::tc_shift_probe_kernel<<<1, 128, 40960, st>>>(ma, mw, K, shift, base_mode, d_out);
  h->launches++;
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(dw);
  if (e != cudaSuccess) return fail(h, MIMI_B200_ERR_CUDA, std::string("debug_shift_probe: ") + cudaGetErrorString(e));
  return MIMI_B200_OK;
}

#ifdef MIMI_TCP_DEBUG
// debug builds only: 64 progress marks of the pair GEMM in mapped host memory (readable while a kernel hangs)
extern "C" unsigned* mimi_b200_debug_marks_init() {
  unsigned* hp = nullptr;
  if (cudaHostAlloc((void**)&hp, 64 * sizeof(unsigned), cudaHostAllocMapped) != cudaSuccess) return nullptr;
  memset(hp, 0, 64 * sizeof(unsigned));
  unsigned* dp = nullptr;
  if (cudaHostGetDevicePointer((void**)&dp, hp, 0) != cudaSuccess) return nullptr;
  if (cudaMemcpyToSymbol(tcp::g_tcp_marks, &dp, sizeof(dp)) != cudaSuccess) return nullptr;
  return hp;
}
#endif

// ---- host staging: a tiny persistent pool of memcpy threads --------------------------------------------------------------
namespace {
struct PackJob { float* dst; const float* src; size_t copy_floats, zero_floats; };
class PackPool {
 public:
  explicit PackPool(int n) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~PackPool() {
    { std::lock_guard<std::mutex> g(m_); stop_ = true; }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // runs the jobs on the workers and the calling thread; returns when all are done
  void run(std::vector<PackJob>& jobs) {
    {
      std::lock_guard<std::mutex> g(m_);
      jobs_ = &jobs; next_ = 0; pending_ = jobs.size(); ++epoch_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [this] { return pending_ == 0; });
    jobs_ = nullptr;
  }

 private:
  // Streaming (non-temporal) stores: the pinned rows are written once and then read by the DMA engine, never by this core,
  // so a regular store would first read every destination line into the cache (read-for-ownership): 3 bytes of host memory
  // traffic per byte staged instead of 2. With 8 ranks staging ~3 GB/s each on one NUMA node that traffic is what bounds the
  // end-to-end rate at 8 GPUs.
  static void stream_copy(float* dst, const float* src, size_t n) {
    size_t i = 0;
    while (i < n && (reinterpret_cast<uintptr_t>(dst + i) & 15)) { dst[i] = src ? src[i] : 0.f; ++i; }
    if (src) {
      for (; i + 16 <= n; i += 16) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 4));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 8));
        const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i + 12));
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 4), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 8), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i + 12), d);
      }
    } else {
      const __m128i z = _mm_setzero_si128();
      for (; i + 4 <= n; i += 4) _mm_stream_si128(reinterpret_cast<__m128i*>(dst + i), z);
    }
    for (; i < n; ++i) dst[i] = src ? src[i] : 0.f;
  }
  static void exec(const PackJob& j) {
    if (j.copy_floats) stream_copy(j.dst, j.src, j.copy_floats);
    if (j.zero_floats) stream_copy(j.dst + j.copy_floats, nullptr, j.zero_floats);
    _mm_sfence();                     // the streamed lines are globally visible before the job counts as done
  }
  void work() {
    for (;;) {
      PackJob j;
      {
        std::lock_guard<std::mutex> g(m_);
        if (!jobs_ || next_ >= jobs_->size()) return;
        j = (*jobs_)[next_++];
      }
      exec(j);
      std::lock_guard<std::mutex> g(m_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
      }
      work();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::vector<PackJob>* jobs_ = nullptr;
  size_t next_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};
std::mutex g_pack_mutex;
std::unique_ptr<PackPool> g_pack_pool;
}  // namespace

int mimi_b200_host_pack(float* h_dst, int64_t dst_stride, const float* const* h_src, const int64_t* h_len,
                        const int64_t* h_zero_to, int n, int n_threads) {
  if (n < 0 || (n > 0 && (!h_dst || !h_src || !h_len || !h_zero_to))) return MIMI_B200_ERR_ARG;
  constexpr size_t kPiece = 1 << 18;                       // 1 MB pieces: a long clip is shared by several threads
  std::vector<PackJob> jobs;
  for (int i = 0; i < n; ++i) {
    if (h_len[i] < 0 || h_len[i] > dst_stride || h_zero_to[i] > dst_stride || (h_len[i] > 0 && !h_src[i])) return MIMI_B200_ERR_ARG;
    float* row = h_dst + (size_t)i * dst_stride;
    const size_t len = (size_t)h_len[i], end = (size_t)std::max<int64_t>(h_len[i], h_zero_to[i]);
    for (size_t o = 0; o < end; o += kPiece) {
      const size_t hi = std::min(end, o + kPiece);
      const size_t c = o < len ? std::min(len, hi) - o : 0;
      jobs.push_back({row + o, h_src[i] ? h_src[i] + o : nullptr, c, hi - o - c});
    }
  }
  if (jobs.empty()) return MIMI_B200_OK;
  std::lock_guard<std::mutex> g(g_pack_mutex);            // one pack at a time per process
  const int want = std::max(1, std::min(n_threads, 16)) - 1;           // the caller is a worker too
  if (!g_pack_pool || g_pack_pool->size() != want) g_pack_pool.reset(new PackPool(want));
  g_pack_pool->run(jobs);
  return MIMI_B200_OK;
}

int64_t mimi_b200_resample_out_len(int64_t n_in, int sr_in, int sr_out) {
  if (sr_in <= 0 || sr_out <= 0 || n_in < 0) return -1;
  if (sr_in == sr_out) return n_in;
  return (int64_t)std::ceil((double)n_in * (double)sr_out / (double)sr_in);
}

int mimi_b200_resample(mimi_b200_t* h, const float* d_in, int64_t in_stride, const int64_t* h_len, int B, int sr_in,
                       int sr_out, float* d_out, int64_t out_stride, void* stream) {
  if (!h || !d_in || !d_out || !h_len) return fail(h, MIMI_B200_ERR_ARG, "resample: NULL argument");
  if (B <= 0) return MIMI_B200_OK;
  if (sr_in <= 0 || sr_out <= 0) return fail(h, MIMI_B200_ERR_ARG, "resample: bad sample rate");
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  std::vector<int> v((size_t)2 * B);
  for (int b = 0; b < B; ++b) {
    if (h_len[b] < 0 || h_len[b] > in_stride || h_len[b] > (1ll << 30)) return fail(h, MIMI_B200_ERR_ARG, "resample: bad length");
    const int64_t ol = mimi_b200_resample_out_len(h_len[b], sr_in, sr_out);
    if (ol > out_stride) return fail(h, MIMI_B200_ERR_ARG, "resample: out_stride too small");
    v[b] = (int)h_len[b];
    v[(size_t)B + b] = (int)ol;
  }
  mimi_b200::Taps tp;
  if (sr_in == sr_out) {
    // identity filter: L = M = 1, single tap 1.0 (REF/*/utils.py:85-86 returns the input unchanged)
    sr_in = sr_out = 1;
  }
  const unsigned long long key = ((unsigned long long)(unsigned)sr_in << 32) | (unsigned)sr_out;
  auto it = h->taps.find(key);
  if (it == h->taps.end()) {
    std::vector<float> taps;
    int c, L, M;
    if (sr_in == sr_out) { taps.assign(1, 1.0f); c = 0; L = 1; M = 1; }
    else design_taps(sr_in, sr_out, taps, c, L, M);
    float* d = nullptr;
    CUDA_TRY(h, cudaMalloc((void**)&d, taps.size() * sizeof(float)));
    CUDA_TRY(h, cudaMemcpy(d, taps.data(), taps.size() * sizeof(float), cudaMemcpyHostToDevice));
    tp = {d, c, L, M, nullptr, 0, 0};
    if (resample_poly_supported(L, M)) {
      // polyphase table g[r][v + Vh][phi] = h[c + phi*M - (v*M + r)*L] (zero outside the prototype), V padded to 8
      const int Vh = (c + L * M + L * M - 1) / (L * M);
      const int V = (2 * Vh + 1 + rsp::kQ - 1) / rsp::kQ * rsp::kQ;
      std::vector<float> g((size_t)M * V * 4, 0.f);
      for (int r = 0; r < M; ++r)
        for (int vv = 0; vv < V; ++vv)
          for (int phi = 0; phi < L; ++phi) {
            const long long idx = (long long)c + (long long)phi * M - ((long long)(vv - Vh) * M + r) * L;
            if (idx >= 0 && idx <= 2ll * c) g[((size_t)r * V + vv) * 4 + phi] = taps[(size_t)idx];
          }
      CUDA_TRY(h, cudaMalloc((void**)&tp.g, g.size() * sizeof(float)));
      CUDA_TRY(h, cudaMemcpy(tp.g, g.data(), g.size() * sizeof(float), cudaMemcpyHostToDevice));
      tp.V = V; tp.Vh = Vh;
    }
    h->taps[key] = tp;
  } else {
    tp = it->second;
  }
  int rc;
  if ((rc = ensure_dev_ints(h, v.size()))) return rc;
  CUDA_TRY(h, cudaStreamWaitEvent(st, h->dev_ints_ev, 0));      // the previous reader of the table (any stream) is done
  if ((rc = stage_ints(h, v, h->dev_ints, st))) return rc;
  if (tp.g && !h->exp_resample_simple) {
    const long long qblocks = (out_stride + tp.L - 1) / tp.L;
    dim3 grid((unsigned)((qblocks + rsp::kQB - 1) / rsp::kQB), B);
    const size_t smem = rsp::smem_floats(tp.L, tp.M, tp.V) * sizeof(float);
#define MIMI_RSP(LL, MM)                                                                                                  \
  if (tp.L == LL && tp.M == MM)                                                                                           \
    resample_poly_kernel<LL, MM><<<grid, rsp::kThreads, smem, st>>>(d_in, in_stride, h->dev_ints, h->dev_ints + B, tp.g, tp.V, \
                                                                     tp.Vh, d_out, out_stride);
    MIMI_RSP(3, 2) MIMI_RSP(1, 2) MIMI_RSP(3, 1) MIMI_RSP(3, 4) MIMI_RSP(2, 1) MIMI_RSP(1, 4) MIMI_RSP(1, 1)
#undef MIMI_RSP
  } else {
    dim3 grid((unsigned)((out_stride + 255) / 256), B);
    resample_kernel<<<grid, 256, 0, st>>>(d_in, in_stride, h->dev_ints, h->dev_ints + B, tp.d, tp.c, tp.L, tp.M, d_out, out_stride);
  }
  h->launches++;
  CUDA_TRY(h, cudaEventRecord(h->dev_ints_ev, st));
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}

int64_t mimi_b200_utf8_bytes_per_frame(int K, uint32_t off, int cbs) {
  if (K < 1 || K > 32 || cbs < 1) return -1;
  const unsigned long long lower = off, upper = (unsigned long long)off + (unsigned long long)K * cbs;
  if (lower < 0xDFFFull && upper > 0xD800ull) return -1;     // converter.py:68-81
  if (upper > 0x110000ull) return -1;
  int64_t n = 0;
  for (int k = 0; k < K; ++k) {
    const unsigned lo = off + (unsigned)k * cbs, hi = lo + cbs - 1;
    if (utf8_len(lo) != utf8_len(hi)) return -1;             // variable width inside one codebook: unsupported
    n += utf8_len(lo);
  }
  return n;
}

int mimi_b200_codes_to_utf8(mimi_b200_t* h, const int64_t* d_codes, int B, int K, int64_t T, const int64_t* h_frames,
                            uint32_t off, int cbs, uint8_t* d_out, int64_t out_stride, int64_t* h_out_len_opt, void* stream) {
  if (!h) return MIMI_B200_ERR_ARG;
  const int64_t bpf = mimi_b200_utf8_bytes_per_frame(K, off, cbs);
  if (bpf < 0) return fail(h, MIMI_B200_ERR_ARG, "codes_to_utf8: unicode offset / codebook range not representable");
  if (B <= 0 || T < 0) return fail(h, MIMI_B200_ERR_ARG, "codes_to_utf8: bad B/T");
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Utf8Params p{};
  p.codes = reinterpret_cast<const long long*>(d_codes); p.B = B; p.K = K; p.T = T; p.offset = off; p.codebook_size = cbs;
  p.bytes_per_frame = (int)bpf; p.out = d_out; p.out_stride = out_stride; p.frames = nullptr;
  int64_t maxfr = T;
  if (h_frames) {
    std::vector<int> v(B);
    maxfr = 0;
    for (int b = 0; b < B; ++b) {
      if (h_frames[b] < 0 || h_frames[b] > T) return fail(h, MIMI_B200_ERR_ARG, "codes_to_utf8: frames out of range");
      v[b] = (int)h_frames[b];
      maxfr = std::max<int64_t>(maxfr, h_frames[b]);
    }
    int rc;
    if ((rc = ensure_dev_ints(h, v.size()))) return rc;
    CUDA_TRY(h, cudaStreamWaitEvent(st, h->dev_ints_ev, 0));
    if ((rc = stage_ints(h, v, h->dev_ints, st))) return rc;
    p.frames = h->dev_ints;
  }
  if (maxfr * bpf > out_stride) return fail(h, MIMI_B200_ERR_ARG, "codes_to_utf8: out_stride too small");
  if (h_out_len_opt)
    for (int b = 0; b < B; ++b) h_out_len_opt[b] = (h_frames ? h_frames[b] : T) * bpf;
  if (maxfr == 0) return MIMI_B200_OK;
  if (!d_codes || !d_out) return fail(h, MIMI_B200_ERR_ARG, "codes_to_utf8: NULL device pointer");
  dim3 grid((unsigned)((maxfr + kUtf8Frames - 1) / kUtf8Frames), B);
  codes_to_utf8_kernel<<<grid, kUtf8Frames, (size_t)kUtf8Frames * bpf + 32, st>>>(p);
  h->launches++;
  if (h_frames) CUDA_TRY(h, cudaEventRecord(h->dev_ints_ev, st));
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}

int mimi_b200_load_decoder_weights(mimi_b200_t* h, const mimi_b200_decoder_weights_t* w) {
  if (!h || !w) return fail(h, MIMI_B200_ERR_ARG, "load_decoder_weights: NULL argument");
  DeviceGuard guard;
  CUDA_TRY(h, cudaSetDevice(h->device));
  return dec_load_weights(h, w);
}

int mimi_b200_decode_workspace_bytes(mimi_b200_t* h, int B, int64_t T, size_t* out_bytes) {
  if (!h || !out_bytes) return fail(h, MIMI_B200_ERR_ARG, "decode_workspace_bytes: NULL argument");
  if (B < 0 || T < 0 || 2 * T > kRopeMaxPos) return fail(h, MIMI_B200_ERR_ARG, "decode_workspace_bytes: bad B / T");
  *out_bytes = make_plan_dec(B, T).bytes + 256;
  return MIMI_B200_OK;
}

int mimi_b200_decode(mimi_b200_t* h, const int64_t* d_codes, int B, int K, int64_t T, float* d_audio, void* d_workspace,
                     size_t workspace_bytes, void* stream) {
  if (!h) return MIMI_B200_ERR_ARG;
  if (!h->loaded || !h->dec_loaded) return fail(h, MIMI_B200_ERR_STATE, "decode: encoder codebooks and decoder weights must be loaded first");
  if (K < 1 || K > MIMI_B200_MAX_QUANTIZERS) return fail(h, MIMI_B200_ERR_ARG, "decode: audio_codes must hold 1..32 codebooks");
  if (B < 0 || T < 0 || 2 * T > kRopeMaxPos) return fail(h, MIMI_B200_ERR_ARG, "decode: bad B / T");
  if (B == 0 || T == 0) return MIMI_B200_OK;
  if (!d_codes || !d_audio || !d_workspace) return fail(h, MIMI_B200_ERR_ARG, "decode: NULL device pointer");
  CUDA_TRY(h, cudaSetDevice(h->device));
  return decode_impl(h, d_codes, B, K, T, d_audio, d_workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mimi_b200_codes_pack_u16(mimi_b200_t* h, const int64_t* d_codes, int64_t n, uint16_t* d_out, void* stream) {
  if (!h) return MIMI_B200_ERR_ARG;
  if (n < 0) return fail(h, MIMI_B200_ERR_ARG, "codes_pack_u16: bad count");
  if (n == 0) return MIMI_B200_OK;
  if (!d_codes || !d_out) return fail(h, MIMI_B200_ERR_ARG, "codes_pack_u16: NULL device pointer");
  CUDA_TRY(h, cudaSetDevice(h->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long quads = (n + 3) / 4;
  const unsigned blocks = (unsigned)std::min<long long>((quads + 255) / 256, 148 * 16);
  codes_pack_u16_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(d_codes), n, d_out);
  h->launches++;
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}

}  // extern "C"
