// Host side of the tensor-core generations (included by mimi_b200.cu after the handle definition): TMA tensor maps, the
// split weight packs (TF32 hi/lo + bf16 for mode 7, row-scaled fp16 hi/lo/hs for mode 9), the split-buffer workspace plan and the
// encode pipeline: fused front end -> CTA-pair tcgen05 GEMM for every other conv / linear -> tcgen05 attention and RVQ.

static int tc_init_driver(mimi_b200* h) {
  if (h->encode_tiled) return MIMI_B200_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CUDA_TRY(h, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn || q != cudaDriverEntryPointSuccess) return fail(h, MIMI_B200_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  h->encode_tiled = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return MIMI_B200_OK;
}

// fp32 tensor map, SWIZZLE_128B, inner box = 32 floats (128 B). rank 2: {K, rows}; rank 3: {K, rows, items}.
static int tc_make_map(mimi_b200* h, CUtensorMap* out, const float* base, int rank, const cuuint64_t* dims,
                       const cuuint64_t* strides_bytes, int box_rows) {
  cuuint32_t box[3] = {(cuuint32_t)tc::kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = h->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims,
                               strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, MIMI_B200_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (dims " +
                                           std::to_string(dims[0]) + "," + std::to_string(dims[1]) + " stride " +
                                           std::to_string(strides_bytes[0]) + ")");
  return MIMI_B200_OK;
}

// bf16 tensor map, SWIZZLE_64B, inner box = 32 elements (64 B): the lo parts and bf16(W_hi) of the bf16-lo generation
static int tc_make_map_bf16(mimi_b200* h, CUtensorMap* out, const void* base, int rank, const cuuint64_t* dims,
                            const cuuint64_t* strides_bytes, int box_rows) {
  cuuint32_t box[3] = {(cuuint32_t)tc::kBK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = h->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                               strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, MIMI_B200_ERR_CUDA, "cuTensorMapEncodeTiled (bf16) failed with CUresult " + std::to_string((int)r));
  return MIMI_B200_OK;
}

// 16-bit tensor map, SWIZZLE_128B, inner box = 64 elements (128 B): the fp16 codebooks of rvq_f16.cuh
static int tc_make_map16_sw128(mimi_b200* h, CUtensorMap* out, const void* base, const cuuint64_t* dims,
                               const cuuint64_t* strides_bytes, int box_rows) {
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = h->encode_tiled(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides_bytes, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(h, MIMI_B200_ERR_CUDA, "cuTensorMapEncodeTiled (16-bit, 128B swizzle) failed with CUresult " + std::to_string((int)r));
  return MIMI_B200_OK;
}

// w_nk: host [N][K] K-major. Mode 7: TF32 hi / lo (fp32) and bf16(hi), uploaded with one map per box height (128, 64, 32
// rows: the pair GEMM stages bnp / 2 weight rows per CTA); map_hi / map_lo (box BN) serve the fused front end.
// `bias` / `scale` (host [N] or nullptr): folded with the mode-9 weight unscaling into the per-column affine of the epilogue,
//   out = act(acc * cmul + cadd):  cmul = unscale * scale,  cadd = bias * scale   (scale = LayerScale of o_proj / fc2)
static int tc_make_weight(mimi_b200* h, TcWeight* w, const std::vector<float>& w_nk, int N, int K, const float* bias = nullptr,
                          const float* scale = nullptr) {
  std::vector<float> hi(w_nk.size()), lo(w_nk.size());
  for (size_t i = 0; i < w_nk.size(); ++i) split_tf32(w_nk[i], hi[i], lo[i]);
  int rc;
  if ((rc = dev_upload(h, &w->hi, hi))) return rc;
  if ((rc = dev_upload(h, &w->lo, lo))) return rc;
  w->N = N; w->K = K; w->BN = (N % 128 == 0) ? 128 : (N % 64 == 0) ? 64 : 32;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)N};
  const cuuint64_t strides[1] = {(cuuint64_t)K * sizeof(float)};
  const cuuint64_t strides_b[1] = {(cuuint64_t)K * 2};
  if ((rc = tc_make_map(h, &w->map_hi, w->hi, 2, dims, strides, w->BN))) return rc;
  if ((rc = tc_make_map(h, &w->map_lo, w->lo, 2, dims, strides, w->BN))) return rc;
  const int rows[3] = {128, 64, 32};
  {
    // bf16(W_hi), round to nearest even (W_hi has 11 significant bits, bf16 keeps 8): the B operand of the A_lo * W_hi term
    std::vector<uint16_t> hb(hi.size());
    for (size_t i = 0; i < hi.size(); ++i) {
      uint32_t u;
      memcpy(&u, &hi[i], 4);
      u += 0x7FFFu + ((u >> 16) & 1u);
      hb[i] = (uint16_t)(u >> 16);
    }
    CUDA_TRY(h, cudaMalloc((void**)&w->hib, hb.size() * sizeof(uint16_t)));
    h->allocs.push_back(w->hib);
    CUDA_TRY(h, cudaMemcpy(w->hib, hb.data(), hb.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    for (int r = 0; r < 3; ++r) {
      if (rows[r] > N) continue;
      if ((rc = tc_make_map(h, &w->m_hi[r], w->hi, 2, dims, strides, rows[r]))) return rc;
      if ((rc = tc_make_map(h, &w->m_lo[r], w->lo, 2, dims, strides, rows[r]))) return rc;
      if ((rc = tc_make_map_bf16(h, &w->m_hib[r], w->hib, 2, dims, strides_b, rows[r]))) return rc;
    }
  }
  {
    // fp16 generation (mode 9): per-row power-of-two scaling, then W' = h16 + l16 with s16 = h16 / 2048
    const size_t nk = (size_t)N * K;
    std::vector<uint16_t> f((size_t)3 * nk);
    std::vector<float> ws(N);
    auto bits = [](float v) { const __half hh = __float2half_rn(v); uint16_t u; memcpy(&u, &hh, 2); return u; };
    for (int n = 0; n < N; ++n) {
      float mx = 0.f;
      for (int k = 0; k < K; ++k) mx = std::max(mx, std::fabs(w_nk[(size_t)n * K + k]));
      int e = 0;
      if (mx > 0.f && std::isfinite(mx)) {
        int ex;
        std::frexp(mx, &ex);                 // mx = m * 2^ex, m in [0.5, 1)  ->  mx * 2^(14 - ex) in [2^13, 2^14)
        e = 14 - ex;
      }
      ws[n] = std::ldexp(1.0f, -e);
      for (int k = 0; k < K; ++k) {
        const float v = std::ldexp(w_nk[(size_t)n * K + k], e);
        const float vh = __half2float(__float2half_rn(v));
        f[(size_t)n * K + k] = bits(vh);
        f[nk + (size_t)n * K + k] = bits(v - vh);
        f[2 * nk + (size_t)n * K + k] = bits(vh * (1.0f / 2048.0f));
      }
    }
    CUDA_TRY(h, cudaMalloc((void**)&w->f16, f.size() * sizeof(uint16_t)));
    h->allocs.push_back(w->f16);
    CUDA_TRY(h, cudaMemcpy(w->f16, f.data(), f.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    std::vector<float> cm7(N), cm9(N), ca(N);
    for (int n = 0; n < N; ++n) {
      const float ls = scale ? scale[n] : 1.0f;
      cm7[n] = ls; cm9[n] = ws[n] * ls; ca[n] = (bias ? bias[n] : 0.0f) * ls;
    }
    if ((rc = dev_upload(h, &w->cmul[0], cm7))) return rc;
    if ((rc = dev_upload(h, &w->cmul[1], cm9))) return rc;
    if ((rc = dev_upload(h, &w->cadd, ca))) return rc;
    if (nk <= 8192) { w->h_f16 = f; w->h_cmul9 = cm9; w->h_cadd = ca; }      // R1a / R1b: the front end builds its own image
    for (int part = 0; part < 3; ++part)
      for (int r = 0; r < 3; ++r) {
        if (rows[r] > N) continue;
        if ((rc = tc_make_map_bf16(h, &w->map_f16[part][r], w->f16 + (size_t)part * nk, 2, dims, strides_b, rows[r]))) return rc;
      }
  }
  return MIMI_B200_OK;
}

// conv weight [Cout][Cin][k] -> [Cout][k*Cin] with K index = tau*Cin + ci (K-major rows)
static std::vector<float> pack_conv_nk(const float* w, int cout, int cin, int k) {
  std::vector<float> t((size_t)cout * cin * k);
  for (int co = 0; co < cout; ++co)
    for (int ci = 0; ci < cin; ++ci)
      for (int tau = 0; tau < k; ++tau) t[((size_t)co * k + tau) * cin + ci] = w[((size_t)co * cin + ci) * k + tau];
  return t;
}

// The resident weights of the fp16 front end (front_f16.cuh) exactly as they sit in its shared memory: K-major rows of 128 bytes
// (64 halfs), 16-byte chunk c of row n at chunk position c ^ (n & 7) (SWIZZLE_128B). W1 = R1a [32][tau * 64 + ci] as nine 4 KB
// blocks [tap][hi | lo | hs]; W2 = R1b [64][32] as three 8 KB parts whose rows use their first 64 bytes.
static int f1_make_image(mimi_b200* h, const mimi_b200_weights_t* w) {
  const TcWeight &w1 = h->tc_conv[1], &w2 = h->tc_conv[2];
  if (w1.N != 32 || w1.K != 192 || w2.N != 64 || w2.K != 32 || w1.h_f16.size() != (size_t)3 * 32 * 192 ||
      w2.h_f16.size() != (size_t)3 * 64 * 32)
    return fail(h, MIMI_B200_ERR_STATE, "front end: unexpected R1a / R1b geometry");
  std::vector<uint16_t> img(f1::kWBytes / 2, 0);
  auto put = [&](size_t byte_off, const uint16_t* src, int rows, int ld, int k0, int kcount) {
    for (int n = 0; n < rows; ++n)
      for (int ci = 0; ci < kcount; ++ci) {
        const int chunk = ci / 8;
        const size_t pos = byte_off + (size_t)n * 128 + (size_t)((chunk ^ (n & 7)) * 16) + (size_t)(ci % 8) * 2;
        img[pos / 2] = src[(size_t)n * ld + k0 + ci];
      }
  };
  for (int tau = 0; tau < 3; ++tau)
    for (int part = 0; part < 3; ++part)
      put((size_t)(tau * 3 + part) * f1::kWBlock, w1.h_f16.data() + (size_t)part * 32 * 192, 32, 192, tau * 64, 64);
  for (int part = 0; part < 3; ++part)
    put((size_t)f1::kW1Bytes + (size_t)part * f1::kW2Part, w2.h_f16.data() + (size_t)part * 64 * 32, 64, 32, 0, 32);
  CUDA_TRY(h, cudaMalloc((void**)&h->f1_wimg, f1::kWBytes));      // (load_weights freed every earlier allocation)
  h->allocs.push_back(h->f1_wimg);
  CUDA_TRY(h, cudaMemcpy(h->f1_wimg, img.data(), f1::kWBytes, cudaMemcpyHostToDevice));
  std::memcpy(h->f1_consts.w0, w->conv_weight[0], sizeof(float) * 64 * 7);
  std::memcpy(h->f1_consts.b0, w->conv_bias[0], sizeof(float) * 64);
  for (int n = 0; n < 32; ++n) { h->f1_consts.m1[n] = w1.h_cmul9[n]; h->f1_consts.a1[n] = w1.h_cadd[n]; }
  for (int n = 0; n < 64; ++n) { h->f1_consts.m2[n] = w2.h_cmul9[n]; h->f1_consts.a2[n] = w2.h_cadd[n]; }
  return MIMI_B200_OK;
}

static int tc_load_weights(mimi_b200* h, const mimi_b200_weights_t* w) {
  int rc;
  if ((rc = tc_init_driver(h))) return rc;
  for (int i = 1; i < MIMI_B200_NUM_CONVS; ++i) {
    const ConvGeom& g = kConv[i];
    if ((rc = tc_make_weight(h, &h->tc_conv[i], pack_conv_nk(w->conv_weight[i], g.cout, g.cin, g.k), g.cout, g.cin * g.k,
                             w->conv_bias[i]))) return rc;
  }
  for (int l = 0; l < MIMI_B200_NUM_LAYERS; ++l) {
    const mimi_b200_layer_weights_t& s = w->layer[l];
    std::vector<float> qkv((size_t)1536 * 512);
    std::memcpy(qkv.data(), s.q_proj_weight, sizeof(float) * 512 * 512);
    std::memcpy(qkv.data() + 512 * 512, s.k_proj_weight, sizeof(float) * 512 * 512);
    std::memcpy(qkv.data() + 2 * 512 * 512, s.v_proj_weight, sizeof(float) * 512 * 512);
    if ((rc = tc_make_weight(h, &h->tc_qkv[l], qkv, 1536, 512))) return rc;
    if ((rc = tc_make_weight(h, &h->tc_o[l], std::vector<float>(s.o_proj_weight, s.o_proj_weight + 512 * 512), 512, 512, nullptr,
                             s.self_attn_layer_scale))) return rc;
    if ((rc = tc_make_weight(h, &h->tc_fc1[l], std::vector<float>(s.fc1_weight, s.fc1_weight + 2048 * 512), 2048, 512))) return rc;
    if ((rc = tc_make_weight(h, &h->tc_fc2[l], std::vector<float>(s.fc2_weight, s.fc2_weight + 512 * 2048), 512, 2048, nullptr,
                             s.mlp_layer_scale))) return rc;
  }
  if ((rc = tc_make_weight(h, &h->tc_down, pack_conv_nk(w->downsample_weight, 512, 512, 4), 512, 2048))) return rc;
  std::vector<float> pj((size_t)512 * 512);
  std::memcpy(pj.data(), w->semantic_input_proj_weight, sizeof(float) * 256 * 512);
  std::memcpy(pj.data() + 256 * 512, w->acoustic_input_proj_weight, sizeof(float) * 256 * 512);
  if ((rc = tc_make_weight(h, &h->tc_proj, pj, 512, 512))) return rc;
  return f1_make_image(h, w);
}

// Workspace plan of the tensor-core generations: raw fp32 buffers where a skip or a non-GEMM consumer needs them, and a
// hi / lo pair (with zero halo rows where a conv pads) for every GEMM operand. Mode 7: hi fp32 (TF32-rounded) + lo bf16,
// 6 bytes per element; `f16` (mode 9): hi and lo both fp16, 4 bytes per element. The 24 kHz level stays on chip (front end).
static PlanTC make_plan_tc(int B, long long N, int K, bool f16) {
  PlanTC p;
  p.B = B; p.K = K; p.N = N;
  long long L = N;
  p.rows[0] = (int)L;
  for (int l = 0; l < 5; ++l) { L = (L + kLevelStride[l] - 1) / kLevelStride[l]; p.rows[l + 1] = (int)L; }
  long long off = 0;
  auto take = [&](long long n) { const long long o = off; off += (n + 63) / 64 * 64; return o; };
  auto raw = [&](int level, int C) { return take((long long)B * p.rows[level] * C); };
  auto split = [&](int level, int C, int front, int back) {
    SplitBuf s;
    s.level = level; s.C = C; s.front = front; s.back = back;
    s.item_stride = (long long)(front + p.rows[level] + back) * C;
    const long long n = (long long)B * s.item_stride + 64;
    s.hi = take(f16 ? (n + 1) / 2 : n);
    s.lo = take((n + 1) / 2);
    return s;
  };
  p.s_h1 = split(0, 64, kHalo, kHalo);
  p.d1 = raw(1, 128);  p.s_d1 = split(1, 128, kHalo, kHalo);  p.s_r2 = split(1, 64, 0, 0);  p.s_h2 = split(1, 128, kHalo, kHalo);
  p.d2 = raw(2, 256);  p.s_d2 = split(2, 256, kHalo, kHalo);  p.s_r3 = split(2, 128, 0, 0); p.s_h3 = split(2, 256, kHalo, kHalo);
  p.d3 = raw(3, 512);  p.s_d3 = split(3, 512, kHalo, kHalo);  p.s_r4 = split(3, 256, 0, 0); p.s_h4 = split(3, 512, kHalo, kHalo);
  p.s_d4 = split(4, 1024, kHalo, kHalo);
  p.z = raw(4, 512);   p.s_y = split(4, 512, 0, 0);  p.qkv = raw(4, 1536);  p.s_att = split(4, 512, 0, 0);
  p.s_ffn = split(4, 2048, 0, 0);
  p.s_zp = split(4, 512, 0, 3);          // rows: [z0, z0, z0..z(T-1), z(T-1)] = T + 3
  p.e = raw(5, 512);   p.s_e = split(5, 512, 0, 0);  p.rp = raw(5, 512);
  p.ints = off * (long long)sizeof(float);
  long long tile_ints = 0;                 // compact tile lists of levels 1..5 (ragged calls)
  for (int l = 1; l < 6; ++l) tile_ints += (long long)B * ((p.rows[l] + 127) / 128);
  p.bytes = (size_t)p.ints + sizeof(int) * (size_t)(7 * B + 1 + 64 + tile_ints);
  return p;
}

namespace {
struct TcCtx {
  mimi_b200* h;
  const PlanTC* p;
  float* ws;
  int B;
  cudaStream_t st;
  const int* const* dlen;     // [6] device length arrays (or nullptr entries)
  const int* maxlen;          // [6]
  std::vector<CUtensorMap>* maps;   // 2 maps (hi, lo) per GEMM site, indexed by a fixed slot id
  uint64_t* built;                  // bit `slot` set once that site's maps are encoded
};
constexpr int kTcSlots = 24;          // GEMM sites; slots [kTcSlots, 2*kTcSlots) hold the flattened-rows maps
}  // namespace

// activation maps for a conv/linear that reads SplitBuf `a` with kernel k, stride s, left pad `pad`. `flat`: the
// items' rows are one contiguous [B * rows][C] matrix (k = 1, no halo) seen as a single item.
// `extra`: box rows and row dimension grow by this many rows (tap-group launches, tc_gemm7.cuh).
static int tc_amaps(TcCtx& c, int slot, const SplitBuf& a, int k, int s, int pad, bool flat, const CUtensorMap** hi,
                    const CUtensorMap** lo, int extra = 0) {
  if (flat) slot += kTcSlots;
  CUtensorMap* mh = &(*c.maps)[2 * slot];
  CUtensorMap* ml = mh + 1;
  if (!((*c.built >> slot) & 1ull)) {
    const PlanTC& p = *c.p;
    const int rows_out = (p.rows[a.level] + s - 1) / s;
    const cuuint64_t dims[3] = {(cuuint64_t)k * a.C, (cuuint64_t)std::max(flat ? rows_out * c.B : rows_out, 1) + extra,
                                (cuuint64_t)(flat ? 1 : c.B)};
    const cuuint64_t strides[2] = {(cuuint64_t)s * a.C * sizeof(float),
                                   (cuuint64_t)a.item_stride * (flat ? c.B : 1) * sizeof(float)};
    const cuuint64_t strides_h[2] = {strides[0] / 2, strides[1] / 2};     // 16-bit arrays with the element indexing of hi
    const long long base_off = (long long)(a.front - pad) * a.C;
    if (a.front < pad) return fail(c.h, MIMI_B200_ERR_ARG, "tc: halo smaller than conv padding");
    int rc;
    if (c.h->mode == 9) {
      if ((rc = tc_make_map_bf16(c.h, mh, reinterpret_cast<const uint16_t*>(c.ws + a.hi) + base_off, 3, dims, strides_h, tc::kBM + extra))) return rc;
    } else if ((rc = tc_make_map(c.h, mh, c.ws + a.hi + base_off, 3, dims, strides, tc::kBM))) return rc;
    if ((rc = tc_make_map_bf16(c.h, ml, reinterpret_cast<const uint16_t*>(c.ws + a.lo) + base_off, 3, dims, strides_h, tc::kBM + extra))) return rc;
    *c.built |= 1ull << slot;
  }
  *hi = mh;
  *lo = ml;
  return MIMI_B200_OK;
}

struct TcOut {
  float* raw = nullptr;            // raw output buffer (rows of N), item stride = rows[level]*N
  const float* res = nullptr;      // residual (same geometry as raw)
  long long raw_item_stride = 0;
  const SplitBuf* split = nullptr; // split output
  int elu_split = 0;
  int act = 0;                     // (bias and LayerScale live in the weight's per-column affine)
};

// k-block order of a conv with k taps, stride s over C_in channels (tc2::Sched): taps grouped by tau mod s
static void tc_korder(const mimi_b200* h, tc2::Sched& sc, int k, int s, int cin) {
  sc.G = 0; sc.s = 0; sc.cp = 0;
  if (h->exp_linear_k == 1 || s <= 1 || k <= s || k % s || cin % 32) return;   // s = 1 (k = 3): taps are 1 row apart, L2 hits anyway
  sc.G = k / s; sc.s = s; sc.cp = cin / 32;
}

// the CTA-pair GEMM (tc_gemm5.cuh): pair tiles of 2 x 128 rows x BNP columns. Every layer of the network has N % 64 == 0.
// Taps per group of a conv the tap-group kernel (tc_gemm7.cuh) can run: 2 for k = 2 s, 3 for k = 3 with s = 1; 1 = tc_gemm5.
static int tc_tap_group(const mimi_b200* h, const TcWeight& w, int k, int s, int cin) {
  if (h->mode != 9 || h->exp_no_taps || k <= 1 || k % s || cin % 32) return 1;
  if (w.N % 256 == 0 && (w.K > 2048 || h->exp_pair_n128 == 1)) return 1;      // 256-column pair tiles: no room for the wider stage
  const int g = k / s;
  return (g == 2 || (g == 3 && s == 1)) ? g : 1;
}

static int launch_tcp(mimi_b200* h, const CUtensorMap& ahi, const CUtensorMap& alo, const TcWeight& w, const tc::Epilogue& ep,
                      int B, int mt_max, cudaStream_t st, int k = 1, int s = 1, int cin = 0, const int* tiles = nullptr,
                      int ntiles = 0, int tg = 1) {
  if (w.N % 64) return fail(h, MIMI_B200_ERR_ARG, "tc: the pair GEMM needs N % 64 == 0");
  // Pair tiles of 256 columns for the deep layers (K > 2048: D3, D4, F -- the tensor pipe is the bound and a 256-wide tile reads
  // every operand byte once per 256 x 256 MMA), 128 columns for everything else: those layers are bound by the tile finish, and
  // a 128-column tile leaves TMEM room for four accumulator chunks, so the MMAs of the next tile run under the finish of the
  // previous one (measured on C2: fc1 -17 %, QKV -9 %, o_proj -15 %, fc2 -8 %, D1 -13 %; D3 / D4 +6 / +11 % if forced to 128).
  // debug_set key 9 = 1 restores 256 columns wherever N allows.
  const bool wide = w.N % 256 == 0 && (w.K > 2048 || h->exp_pair_n128 == 1);
  const int bnp = wide ? 256 : (w.N % 128 == 0) ? 128 : 64;
  tcp::Sched sc{B, mt_max, w.N / bnp};
  tc_korder(h, sc, k, s, cin);
  sc.tiles = tiles; sc.ntiles = ntiles;
  const long long npairs = (((tiles ? (long long)ntiles : (long long)mt_max * B) + 1) / 2) * sc.ntn;
  const int ncl = (int)std::min<long long>(npairs, h->num_clusters);
  if (ncl <= 0) return MIMI_B200_OK;
  const int r = bnp == 256 ? 0 : bnp == 128 ? 1 : 2;            // weight boxes of bnp / 2 rows
  if (tg > 1) {
    // taps that share input rows: one activation tile of 128 + tg - 1 rows per (tap phase, channel panel) (tc_gemm7.cuh)
    sc.G = tg; sc.s = s; sc.cp = cin / 32;
    if (w.K != tg * sc.s * sc.cp * 32 || bnp == 256) return fail(h, MIMI_B200_ERR_ARG, "tc: tap-group launch with a wrong geometry");
    const CUtensorMap &whi = w.map_f16[0][r], &wlo = w.map_f16[1][r], &whs = w.map_f16[2][r];
    if (bnp == 128 && tg == 2) tcg::tcp_taps_kernel<128, 2><<<2 * ncl, tcg::kThreads, tcg::Cfg<128, 2>::SMEM, st>>>(ahi, alo, whi, wlo, whs, ep, sc);
    else if (bnp == 128 && tg == 3) tcg::tcp_taps_kernel<128, 3><<<2 * ncl, tcg::kThreads, tcg::Cfg<128, 3>::SMEM, st>>>(ahi, alo, whi, wlo, whs, ep, sc);
    else if (bnp == 64 && tg == 3) tcg::tcp_taps_kernel<64, 3><<<2 * ncl, tcg::kThreads, tcg::Cfg<64, 3>::SMEM, st>>>(ahi, alo, whi, wlo, whs, ep, sc);
    else if (bnp == 64 && tg == 2) tcg::tcp_taps_kernel<64, 2><<<2 * ncl, tcg::kThreads, tcg::Cfg<64, 2>::SMEM, st>>>(ahi, alo, whi, wlo, whs, ep, sc);
    return MIMI_B200_OK;
  }
  if (h->mode == 9) {
    // fp16 generation: all five operand tiles are 16-bit SWIZZLE_64B boxes
    const CUtensorMap &whi = w.map_f16[0][r], &wlo = w.map_f16[1][r], &whs = w.map_f16[2][r];
    if (bnp == 256) tcp::tcp_gemm_kernel<256, 3><<<2 * ncl, tcp::kThreads, tcp::Cfg<256, 3>::SMEM, st>>>(ahi, alo, whi, wlo, whs, w.K, ep, sc);
    else if (bnp == 128) tcp::tcp_gemm_kernel<128, 3><<<2 * ncl, tcp::kThreads, tcp::Cfg<128, 3>::SMEM, st>>>(ahi, alo, whi, wlo, whs, w.K, ep, sc);
    else tcp::tcp_gemm_kernel<64, 3><<<2 * ncl, tcp::kThreads, tcp::Cfg<64, 3>::SMEM, st>>>(ahi, alo, whi, wlo, whs, w.K, ep, sc);
  } else {
    // TF32 generation with bf16 lo parts (mode 7): alo is a bf16 map, the third weight tile is bf16(W_hi)
    const CUtensorMap &whi = w.m_hi[r], &wlo = w.m_lo[r], &whb = w.m_hib[r];
    if (bnp == 256) tcp::tcp_gemm_kernel<256, 1><<<2 * ncl, tcp::kThreads, tcp::Cfg<256, 1>::SMEM, st>>>(ahi, alo, whi, wlo, whb, w.K, ep, sc);
    else if (bnp == 128) tcp::tcp_gemm_kernel<128, 1><<<2 * ncl, tcp::kThreads, tcp::Cfg<128, 1>::SMEM, st>>>(ahi, alo, whi, wlo, whb, w.K, ep, sc);
    else tcp::tcp_gemm_kernel<64, 1><<<2 * ncl, tcp::kThreads, tcp::Cfg<64, 1>::SMEM, st>>>(ahi, alo, whi, wlo, whb, w.K, ep, sc);
  }
  return MIMI_B200_OK;
}

static int tc_gemm(TcCtx& c, int slot, const SplitBuf& a, int k, int s, int pad, const TcWeight& w, const TcOut& o, int prof_id) {
  const CUtensorMap *ahi = nullptr, *alo = nullptr;
  int rc;
  if (slot < 0 || slot >= kTcSlots) return fail(c.h, MIMI_B200_ERR_ARG, "tc: bad map slot");
  // Linears (k = 1) over buffers without halo rows: the B items are one contiguous [B * rows][C] matrix. When whole
  // 128-row tiles of that matrix are fewer than the per-item tiles (each item rounds up on its own), run it as ONE item:
  // rows past an item's length are computed but not stored (every consumer is row-wise and length-bound).
  const int rows_lvl = c.p->rows[a.level];
  const bool flat = !c.h->exp_no_flat && k == 1 && s == 1 && pad == 0 && a.front == 0 && a.back == 0 &&
                    c.B > 1 && (!o.raw && !o.res || o.raw_item_stride == (long long)rows_lvl * w.N) &&
                    (!o.split || (o.split->front == 0 && o.split->back == 0 && o.split->level == a.level)) &&
                    ((long long)c.B * rows_lvl + 127) / 128 < c.h->item_tiles[a.level];
  const int tg = flat ? 1 : tc_tap_group(c.h, w, k, s, a.C);
  if ((rc = tc_amaps(c, slot, a, k, s, pad, flat, &ahi, &alo, tg - 1))) return rc;
  if (w.K != k * a.C) return fail(c.h, MIMI_B200_ERR_ARG, "tc: weight K mismatch");
  tc::Epilogue ep{};
  ep.cmul = w.cmul[c.h->mode == 9 ? 1 : 0]; ep.cadd = w.cadd;
  ep.res = o.res; ep.out_raw = o.raw; ep.raw_item_stride = o.raw_item_stride;
  if (o.split) {
    if (o.split->C != w.N) return fail(c.h, MIMI_B200_ERR_ARG, "tc: split output width mismatch");
    ep.out_hi = c.ws + o.split->hi; ep.out_lo = c.ws + o.split->lo;
    ep.split_item_stride = o.split->item_stride; ep.split_front = o.split->front;
  }
  ep.act = (o.act == 1 && !c.h->exp_gelu_erff) ? 2 : o.act; ep.elu_split = o.elu_split; ep.lo_bf16 = c.h->mode == 9 ? 3 : 1;
  ep.chunk_kb = c.h->exp_chunk_kb;
  ep.len_in = c.dlen[a.level]; ep.uniform_len_in = c.maxlen[a.level]; ep.conv_stride = s; ep.N = w.N;
  int lout_max = (c.maxlen[a.level] + s - 1) / s;
  if (lout_max <= 0) return MIMI_B200_OK;
  int nb = c.B;
  if (flat) {
    ep.flat_len = c.dlen[a.level]; ep.flat_rows = rows_lvl;      // rows past an item's length: computed, not stored
    ep.len_in = nullptr; ep.uniform_len_in = c.B * rows_lvl; lout_max = c.B * rows_lvl; nb = 1;
  }
  // ragged call: the compact list of the output level's tiles (the output of a strided conv lives one level up)
  const int out_level = a.level + (s > 1 ? 1 : 0);
  const int* tiles = (!flat && out_level < 6) ? c.h->tile_ptr[out_level] : nullptr;
  const int ntiles = tiles ? c.h->tile_cnt[out_level] : 0;
  if ((rc = launch_tcp(c.h, *ahi, *alo, w, ep, nb, (lout_max + tc::kBM - 1) / tc::kBM, c.st, k, s, a.C, tiles, ntiles, tg))) return rc;
  c.h->launches++;
  mark(c.h, prof_id, c.st);
  CUDA_TRY(c.h, cudaGetLastError());
  return MIMI_B200_OK;
}

static int tc_zero_halo(TcCtx& c, const SplitBuf& s) {
  if (s.front + s.back == 0 || c.B == 0) return MIMI_B200_OK;
  const int per = (s.front + s.back) * s.C;
  dim3 grid((per + 255) / 256, c.B);
  tc::zero_halo_kernel<<<grid, 256, 0, c.st>>>(c.ws + s.hi, c.ws + s.lo, s.item_stride, s.C, s.front, s.back,
                                               c.dlen[s.level], c.maxlen[s.level], c.h->mode == 9 ? 3 : 1);
  c.h->launches++;
  mark(c.h, 25, c.st);
  return MIMI_B200_OK;
}

// The encode pipeline of the tensor-core generations (modes 7 and 9). Same contract as the fp32 FFMA pipeline in mimi_b200.cu.
static int encode_tc(mimi_b200* h, const float* d_input, int B, long long N, int K, const PlanTC& p, float* ws,
                     const int* const* dlen, const int* maxlen, const int* dprefix, int total_frames,
                     int64_t* d_codes, float* d_latent_opt, cudaStream_t st) {
  int rc;
  const int lob = h->mode == 9 ? 3 : 1;            // split format every producer writes (common.cuh: store_split4_x)
  const MapKey key{ws, B, N, h->mode};             // one set of activation maps per (workspace, shape, generation)
  auto it = h->amap_cache.find(key);
  TcCtx c{h, &p, ws, B, st, dlen, maxlen, nullptr, nullptr};
  if (it == h->amap_cache.end()) {
    if (h->amap_cache.size() >= 64) h->amap_cache.clear();
    it = h->amap_cache.emplace(key, MapSet()).first;
    it->second.maps.resize(4 * kTcSlots);
  }
  c.maps = &it->second.maps;
  c.built = &it->second.built;
  auto rstride = [&](int level, int C) { return (long long)p.rows[level] * C; };

  // halo rows of every conv-consumed split buffer (producers only ever write rows [0, L))
  const SplitBuf* halos[] = {&p.s_h1, &p.s_d1, &p.s_h2, &p.s_d2, &p.s_h3, &p.s_d3, &p.s_h4, &p.s_d4};
  if (h->phase <= MIMI_B200_PHASE_BEGIN)
    for (const SplitBuf* s : halos)
      if ((rc = tc_zero_halo(c, *s))) return rc;
  if (h->phase == MIMI_B200_PHASE_BEGIN) return MIMI_B200_OK;

  // ---- fused front end: waveform -> L0 -> R1a -> R1b (+skip) -> ELU -> split; the 24 kHz activations stay on chip ------------
  // Items [b0, b1) only in a phased call (every item is independent here); PHASE_FINISH finds the front end already done.
  if (h->phase != MIMI_B200_PHASE_FINISH) {
    const int b0 = h->phase == MIMI_B200_PHASE_FRONT ? h->front_b0 : 0;
    const int nb = (h->phase == MIMI_B200_PHASE_FRONT ? h->front_b1 : B) - b0;
    if (maxlen[0] > 0 && nb > 0) {
      f0::Params fp{};
      fp.x = d_input + (long long)b0 * N; fp.x_stride = N; fp.len = dlen[0] ? dlen[0] + b0 : nullptr; fp.uniform_len = maxlen[0];
      fp.B = nb;
      fp.mt_max = (maxlen[0] + f0::kAdv - 1) / f0::kAdv;
      // (16-bit arrays: the same ELEMENT offset is half as many floats)
      fp.out_hi = ws + p.s_h1.hi + (long long)b0 * p.s_h1.item_stride / (h->mode == 9 ? 2 : 1);
      fp.out_lo = ws + p.s_h1.lo + (long long)b0 * p.s_h1.item_stride / 2;
      fp.split_item_stride = p.s_h1.item_stride; fp.split_front = p.s_h1.front;
      fp.lo_bf16 = lob;
      const long long vt = (long long)fp.mt_max * nb;
      const int grid = (int)std::min<long long>(vt, h->num_sms);
      if (h->mode == 9 && !h->exp_front_tf32) {
        // fp16 generation: front_f16.cuh walks only the tiles that exist, item by item
        f1::Params q{};
        q.x = fp.x; q.x_stride = N; q.len = fp.len; q.uniform_len = fp.uniform_len; q.B = nb;
        long long tiles = 0;
        for (int b = b0; b < b0 + nb; ++b) tiles += (std::min<long long>(h->len0_host[b], N) + f1::kAdv - 1) / f1::kAdv;
        if (tiles > 0x7fffffffll) return fail(h, MIMI_B200_ERR_ARG, "front end: too many tiles");
        q.total_tiles = (int)tiles;
        q.wimg = h->f1_wimg;
        q.out_hi = reinterpret_cast<uint16_t*>(fp.out_hi); q.out_lo = reinterpret_cast<uint16_t*>(fp.out_lo);
        q.split_item_stride = fp.split_item_stride; q.split_front = fp.split_front;
        const int g1 = (int)std::min<long long>((tiles + 1) / 2, h->num_sms);
        if (g1 > 0) f1::front_f16_kernel<<<g1, f1::kThreads, f1::kSmem, st>>>(h->f1_consts, q);
      } else {
        f0::front_fused_kernel<<<grid, f0::kThreads, f0::kSmem, st>>>(h->tc_conv[1].map_hi, h->tc_conv[1].map_lo, h->tc_conv[2].map_hi,
                                                                      h->tc_conv[2].map_lo, h->f0_consts, fp);
      }
      h->launches++; mark(h, 27, st);
      CUDA_TRY(h, cudaGetLastError());
    }
  }
  if (h->phase == MIMI_B200_PHASE_FRONT) return MIMI_B200_OK;

  // ---- D1 .. R4b ------------------------------------------------------------------------------------------------------
  struct Lvl { const SplitBuf* in; long long d_raw; const SplitBuf *s_d, *s_r, *s_h; int C; };   // C = channels after the down conv
  const Lvl lv[3] = {{&p.s_h1, p.d1, &p.s_d1, &p.s_r2, &p.s_h2, 128},
                     {&p.s_h2, p.d2, &p.s_d2, &p.s_r3, &p.s_h3, 256},
                     {&p.s_h3, p.d3, &p.s_d3, &p.s_r4, &p.s_h4, 512}};
  for (int s = 0; s < 3; ++s) {
    const Lvl& L = lv[s];
    const int id = 3 + 3 * s, ia = 4 + 3 * s, ib = 5 + 3 * s;
    const ConvGeom& gd = kConv[id];
    TcOut o;   // down conv: raw (skip) + ELU'd split (resblock conv a)
    o.raw = ws + L.d_raw; o.raw_item_stride = rstride(s + 1, L.C); o.split = L.s_d; o.elu_split = 1;
    if ((rc = tc_gemm(c, id, *L.in, gd.k, gd.stride, gd.k - gd.stride, h->tc_conv[id], o, id))) return rc;
    o = TcOut{};   // resblock conv a: C -> C/2, k3
    o.split = L.s_r; o.elu_split = 1;
    if ((rc = tc_gemm(c, ia, *L.s_d, 3, 1, 2, h->tc_conv[ia], o, ia))) return rc;
    o = TcOut{};   // resblock conv b: C/2 -> C, k1, + skip; only ELU(h) is needed downstream
    o.res = ws + L.d_raw; o.raw_item_stride = rstride(s + 1, L.C); o.split = L.s_h; o.elu_split = 1;
    if ((rc = tc_gemm(c, ib, *L.s_r, 1, 1, 0, h->tc_conv[ib], o, ib))) return rc;
  }
  {
    TcOut o;   // D4: 512 -> 1024, k16 s8
    o.split = &p.s_d4; o.elu_split = 1;
    if ((rc = tc_gemm(c, 12, p.s_h4, 16, 8, 8, h->tc_conv[12], o, 12))) return rc;
    o = TcOut{};   // F: 1024 -> 512, k3 -> transformer stream z (raw)
    o.raw = ws + p.z; o.raw_item_stride = rstride(4, 512);
    if ((rc = tc_gemm(c, 13, p.s_d4, 3, 1, 2, h->tc_conv[13], o, 13))) return rc;
  }
  // ---- encoder transformer -------------------------------------------------------------------------------------
  const int T25 = maxlen[4];
  for (int l = 0; l < h->dbg_layers && T25 > 0; ++l) {
    const LayerDev& d = h->layer[l];
    dim3 lgrid((T25 + 7) / 8, B);
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z, ws + p.s_y.hi, d.ln1_w, d.ln1_b, rstride(4, 512), dlen[4], T25, ws + p.s_y.lo, lob);
    h->launches++; mark(h, 14, st);
    TcOut o;
    o.raw = ws + p.qkv; o.raw_item_stride = rstride(4, 1536);
    if ((rc = tc_gemm(c, 0, p.s_y, 1, 1, 0, h->tc_qkv[l], o, 15))) return rc;
    {
      // tensor-core attention: persistent CTAs over (128-query tile, head, item) units
      atc::Params ap{};
      ap.qkv = ws + p.qkv; ap.item_stride = rstride(4, 1536); ap.out_hi = ws + p.s_att.hi; ap.out_lo = ws + p.s_att.lo;
      ap.out_stride = rstride(4, 512); ap.rope_cos = h->rope_cos; ap.rope_sin = h->rope_sin; ap.len = dlen[4];
      ap.uniform_len = T25; ap.B = B; ap.mt_max = (T25 + atc::kQT - 1) / atc::kQT; ap.lo_bf16 = lob;
      if (!h->exp_att_grid && h->tile_ptr[4]) { ap.tiles = h->tile_ptr[4]; ap.ntiles = h->tile_cnt[4]; }
      const long long units = (ap.tiles ? (long long)ap.ntiles : (long long)ap.mt_max * B) * kHeads;
      atc::swa_attention_tc_kernel<<<(int)std::min<long long>(units, h->num_sms), atc::kThreads, atc::kSmem, st>>>(ap);
      h->launches++; mark(h, 16, st);
      CUDA_TRY(h, cudaGetLastError());
    }
    o = TcOut{};   // o_proj + LayerScale + residual, in place on z
    o.raw = ws + p.z; o.res = ws + p.z; o.raw_item_stride = rstride(4, 512);
    if ((rc = tc_gemm(c, 1, p.s_att, 1, 1, 0, h->tc_o[l], o, 17))) return rc;
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z, ws + p.s_y.hi, d.ln2_w, d.ln2_b, rstride(4, 512), dlen[4], T25, ws + p.s_y.lo, lob);
    h->launches++; mark(h, 14, st);
    o = TcOut{};   // fc1 + GELU(erf) -> split
    o.split = &p.s_ffn; o.act = 1;
    if ((rc = tc_gemm(c, 0, p.s_y, 1, 1, 0, h->tc_fc1[l], o, 18))) return rc;
    o = TcOut{};   // fc2 + LayerScale + residual, in place on z
    o.raw = ws + p.z; o.res = ws + p.z; o.raw_item_stride = rstride(4, 512);
    if ((rc = tc_gemm(c, 2, p.s_ffn, 1, 1, 0, h->tc_fc2[l], o, 19))) return rc;
  }
  // ---- stride-2 downsample (replicate pad materialised as 3 extra rows), RVQ input projections ------------------
  const int T = maxlen[5];
  if (T25 > 0) {
    dim3 pgrid((T25 + 3 + 7) / 8, B);
    tc::pad_replicate_split_kernel<<<pgrid, 256, 0, st>>>(ws + p.z, rstride(4, 512), ws + p.s_zp.hi, ws + p.s_zp.lo,
                                                          p.s_zp.item_stride, dlen[4], T25, lob);
    h->launches++; mark(h, 26, st);
    TcOut o;
    o.raw = ws + p.e; o.raw_item_stride = rstride(5, 512); o.split = &p.s_e;
    // the downsample conv reads s_zp rows 2j .. 2j+3: k=4, stride 2, no further padding. Lout must come from
    // the T25 lengths (ceil(T25/2)), which is what tc_gemm derives from level-4 lengths with s = 2.
    if ((rc = tc_gemm(c, 14, p.s_zp, 4, 2, 0, h->tc_down, o, 20))) return rc;
  }
  if (d_latent_opt && T > 0) {
    const long long n = (long long)kHidden * p.rows[5];
    dim3 tgrid((unsigned)((n + 255) / 256), B);
    latent_transpose_kernel<<<tgrid, 256, 0, st>>>(ws + p.e, rstride(5, 512), d_latent_opt, p.rows[5], dlen[5], T);
    h->launches++; mark(h, 23, st);
  }
  if (T > 0) {
    TcOut o;
    o.raw = ws + p.rp; o.raw_item_stride = rstride(5, 512);
    if ((rc = tc_gemm(c, 15, p.s_e, 1, 1, 0, h->tc_proj, o, 21))) return rc;
    if (total_frames > 0 && h->mode == 9 && !h->exp_rvq_tf32) {
      rvq16::Params q{};
      q.rproj = ws + p.rp; q.item_stride = rstride(5, 512); q.embed = h->embed; q.enorm = h->enorm; q.m2s = h->embed_m2s;
      q.codes = reinterpret_cast<long long*>(d_codes);
      q.K = K; q.T_out = p.rows[5]; q.len = dlen[5]; q.uniform_len = T; q.B = B; q.total_frames = total_frames;
      q.frame_prefix = dprefix;
      rvq16::rvq_f16_kernel<<<(total_frames + rvq16::kFrames - 1) / rvq16::kFrames, rvq16::kThreads, rvq16::kSmem, st>>>(
          h->map_embed16_hi, h->map_embed16_lo, q);
      h->launches++; mark(h, 22, st);
    } else if (total_frames > 0) {
      rvqtc::Params q{};
      q.rproj = ws + p.rp; q.item_stride = rstride(5, 512); q.embed = h->embed; q.enorm = h->enorm;
      q.codes = reinterpret_cast<long long*>(d_codes);
      q.K = K; q.T_out = p.rows[5]; q.len = dlen[5]; q.uniform_len = T; q.B = B; q.total_frames = total_frames;
      q.frame_prefix = dprefix;
      rvqtc::rvq_tc_kernel<<<(total_frames + rvqtc::kFrames - 1) / rvqtc::kFrames, rvqtc::kThreads, rvqtc::kSmem, st>>>(
          h->map_embed_hi, h->map_embed_lo, q);
      h->launches++; mark(h, 22, st);
    }
  }
  if (h->mode == 9 && h->range_flag)    // the fp16 generation's saturation flag rides behind the call on its stream
    CUDA_TRY(h, cudaMemcpyFromSymbolAsync(h->range_flag, g_f16_overflow, sizeof(int), 0, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}
