// Host side of the fourth-generation path (mode 5, included by mimi_b200.cu after tc_host.inl): every activation is
// one fp32 buffer (with zero halo rows where a conv pads), the GEMM kernel (tc_gemm4.cuh) applies the consumer's ELU
// and the TF32 hi/lo split in shared memory.

static PlanR make_plan_r(int B, long long N, int K) {
  PlanR p;
  p.B = B; p.K = K; p.N = N;
  long long L = N;
  p.rows[0] = (int)L;
  for (int l = 0; l < 5; ++l) { L = (L + kLevelStride[l] - 1) / kLevelStride[l]; p.rows[l + 1] = (int)L; }
  long long off = 0;
  auto buf = [&](int level, int C, int front, int back) {
    RawBuf r;
    r.level = level; r.C = C; r.front = front; r.back = back;
    r.item_stride = (long long)(front + p.rows[level] + back) * C;
    r.off = off;
    off += ((long long)B * r.item_stride + 64 + 63) / 64 * 64;
    return r;
  };
  p.h1 = buf(0, 64, kHalo, kHalo);
  p.d1 = buf(1, 128, kHalo, kHalo);  p.r2 = buf(1, 64, 0, 0);   p.h2 = buf(1, 128, kHalo, kHalo);
  p.d2 = buf(2, 256, kHalo, kHalo);  p.r3 = buf(2, 128, 0, 0);  p.h3 = buf(2, 256, kHalo, kHalo);
  p.d3 = buf(3, 512, kHalo, kHalo);  p.r4 = buf(3, 256, 0, 0);  p.h4 = buf(3, 512, kHalo, kHalo);
  p.d4 = buf(4, 1024, kHalo, kHalo);
  p.z = buf(4, 512, 0, 0);  p.y = buf(4, 512, 0, 0);  p.qkv = buf(4, 1536, 0, 0);  p.att = buf(4, 512, 0, 0);
  p.ffn = buf(4, 2048, 0, 0);
  p.zp = buf(4, 512, 0, 3);            // rows: [z0, z0, z0..z(T-1), z(T-1)] = T + 3
  p.e = buf(5, 512, 0, 0);  p.rp = buf(5, 512, 0, 0);
  p.ints = off * (long long)sizeof(float);
  p.bytes = (size_t)p.ints + sizeof(int) * (size_t)(7 * B + 1 + 64);
  return p;
}

namespace {
struct R5Ctx {
  mimi_b200* h;
  const PlanR* p;
  float* ws;
  int B;
  cudaStream_t st;
  const int* const* dlen;
  const int* maxlen;
  std::vector<CUtensorMap>* maps;   // one map per GEMM site
  uint64_t* built;
};
struct R5Out {
  const RawBuf* dst = nullptr;
  const RawBuf* res = nullptr;     // residual buffer (same geometry as dst; may be dst itself)
  const float* bias = nullptr;
  const float* scale = nullptr;
  int act = 0;
};
}  // namespace

static float* r5_rows(const R5Ctx& c, const RawBuf& b) { return c.ws + b.off + (long long)b.front * b.C; }

static int r5_gemm(R5Ctx& c, int slot, const RawBuf& a, int k, int s, int pad, int elu_in, const TcWeight& w, const R5Out& o,
                   int prof_id) {
  if (slot < 0 || slot >= kTcSlots) return fail(c.h, MIMI_B200_ERR_ARG, "tc5: bad map slot");
  if (w.K != k * a.C) return fail(c.h, MIMI_B200_ERR_ARG, "tc5: weight K mismatch");
  if (a.front < pad) return fail(c.h, MIMI_B200_ERR_ARG, "tc5: halo smaller than conv padding");
  if (!o.dst || o.dst->C != w.N) return fail(c.h, MIMI_B200_ERR_ARG, "tc5: output width mismatch");
  CUtensorMap* m = &(*c.maps)[slot];
  int rc;
  if (!((*c.built >> slot) & 1ull)) {
    const int rows_out = (c.p->rows[a.level] + s - 1) / s;
    const cuuint64_t dims[3] = {(cuuint64_t)k * a.C, (cuuint64_t)std::max(rows_out, 1), (cuuint64_t)c.B};
    const cuuint64_t strides[2] = {(cuuint64_t)s * a.C * sizeof(float), (cuuint64_t)a.item_stride * sizeof(float)};
    if ((rc = tc_make_map(c.h, m, c.ws + a.off + (long long)(a.front - pad) * a.C, 3, dims, strides, tc::kBM))) return rc;
    *c.built |= 1ull << slot;
  }
  tc::Epilogue ep{};
  ep.bias = o.bias; ep.scale = o.scale; ep.act = o.act;
  ep.out_raw = r5_rows(c, *o.dst); ep.raw_item_stride = o.dst->item_stride;
  if (o.res) {
    if (o.res->item_stride != o.dst->item_stride || o.res->C != o.dst->C)
      return fail(c.h, MIMI_B200_ERR_ARG, "tc5: residual geometry differs from the output's");
    ep.res = r5_rows(c, *o.res);
  }
  ep.len_in = c.dlen[a.level]; ep.uniform_len_in = c.maxlen[a.level]; ep.conv_stride = s; ep.N = w.N;
  ep.chunk_kb = c.h->exp_chunk_kb;
  const int lout_max = (c.maxlen[a.level] + s - 1) / s;
  if (lout_max <= 0) return MIMI_B200_OK;
  tc2::Sched sc{c.B, (lout_max + tc::kBM - 1) / tc::kBM, w.N / w.BN};
  const long long vt = (long long)sc.mt_max * c.B * sc.ntn;
  const int grid = (int)std::min<long long>(vt, c.h->num_sms);
  if (w.BN == 128)
    tc4::tc4_gemm_kernel<128><<<grid, tc4::kThreads, tc4::Cfg<128>::SMEM, c.st>>>(*m, w.map_hi, w.map_lo, w.K, elu_in, ep, sc);
  else if (w.BN == 64)
    tc4::tc4_gemm_kernel<64><<<grid, tc4::kThreads, tc4::Cfg<64>::SMEM, c.st>>>(*m, w.map_hi, w.map_lo, w.K, elu_in, ep, sc);
  else
    tc4::tc4_gemm_kernel<32><<<grid, tc4::kThreads, tc4::Cfg<32>::SMEM, c.st>>>(*m, w.map_hi, w.map_lo, w.K, elu_in, ep, sc);
  c.h->launches++;
  mark(c.h, prof_id, c.st);
  CUDA_TRY(c.h, cudaGetLastError());
  return MIMI_B200_OK;
}

static int r5_zero_halo(R5Ctx& c, const RawBuf& b) {
  if (b.front + b.back == 0 || c.B == 0) return MIMI_B200_OK;
  const int per = (b.front + b.back) * b.C;
  dim3 grid((per + 255) / 256, c.B);
  tc4::zero_halo_raw_kernel<<<grid, 256, 0, c.st>>>(c.ws + b.off, b.item_stride, b.C, b.front, b.back, c.dlen[b.level], c.maxlen[b.level]);
  c.h->launches++;
  mark(c.h, 25, c.st);
  return MIMI_B200_OK;
}

static int encode_tc5(mimi_b200* h, const float* d_input, int B, long long N, int K, const PlanR& p, float* ws,
                      const int* const* dlen, const int* maxlen, const int* dprefix, int total_frames,
                      int64_t* d_codes, float* d_latent_opt, cudaStream_t st) {
  int rc;
  const MapKey key{ws, B, N, 2};
  auto it = h->amap_cache.find(key);
  R5Ctx c{h, &p, ws, B, st, dlen, maxlen, nullptr, nullptr};
  if (it == h->amap_cache.end()) {
    if (h->amap_cache.size() >= 64) h->amap_cache.clear();
    it = h->amap_cache.emplace(key, MapSet()).first;
    it->second.maps.resize(kTcSlots);
  }
  c.maps = &it->second.maps;
  c.built = &it->second.built;

  const RawBuf* halos[] = {&p.h1, &p.d1, &p.h2, &p.d2, &p.h3, &p.d3, &p.h4, &p.d4};
  for (const RawBuf* b : halos)
    if ((rc = r5_zero_halo(c, *b))) return rc;

  // ---- fused 24 kHz front end: waveform -> L0 -> R1a -> R1b (+skip), raw fp32 out -------------------------------------
  if (maxlen[0] > 0) {
    f0::Params fp{};
    fp.x = d_input; fp.x_stride = N; fp.len = dlen[0]; fp.uniform_len = maxlen[0]; fp.B = B;
    fp.mt_max = (maxlen[0] + f0::kAdv - 1) / f0::kAdv;
    fp.out_hi = ws + p.h1.off; fp.out_lo = nullptr; fp.split_item_stride = p.h1.item_stride; fp.split_front = p.h1.front;
    fp.raw_out = 1;
    const long long vt = (long long)fp.mt_max * B;
    const int grid = (int)std::min<long long>(vt, h->num_sms);
    f0::front_fused_kernel<<<grid, f0::kThreads, f0::kSmem, st>>>(h->tc_conv[1].map_hi, h->tc_conv[1].map_lo, h->tc_conv[2].map_hi,
                                                                  h->tc_conv[2].map_lo, h->f0_consts, fp);
    h->launches++; mark(h, 27, st);
    CUDA_TRY(h, cudaGetLastError());
  }
  // ---- D1 .. R4b, D4, F ---------------------------------------------------------------------------------------------------
  struct Lvl { const RawBuf *in, *d, *r, *hh; };
  const Lvl lv[3] = {{&p.h1, &p.d1, &p.r2, &p.h2}, {&p.h2, &p.d2, &p.r3, &p.h3}, {&p.h3, &p.d3, &p.r4, &p.h4}};
  for (int s = 0; s < 3; ++s) {
    const Lvl& L = lv[s];
    const int id = 3 + 3 * s, ia = 4 + 3 * s, ib = 5 + 3 * s;
    const ConvGeom& gd = kConv[id];
    R5Out o;
    o.dst = L.d; o.bias = h->conv_b[id];                                  // down conv: ELU(h) -> d (raw)
    if ((rc = r5_gemm(c, id, *L.in, gd.k, gd.stride, gd.k - gd.stride, 1, h->tc_conv[id], o, id))) return rc;
    o = R5Out{}; o.dst = L.r; o.bias = h->conv_b[ia];                    // resblock conv a: ELU(d) -> r
    if ((rc = r5_gemm(c, ia, *L.d, 3, 1, 2, 1, h->tc_conv[ia], o, ia))) return rc;
    o = R5Out{}; o.dst = L.hh; o.res = L.d; o.bias = h->conv_b[ib];      // resblock conv b: ELU(r) -> h = d + conv
    if ((rc = r5_gemm(c, ib, *L.r, 1, 1, 0, 1, h->tc_conv[ib], o, ib))) return rc;
  }
  {
    R5Out o; o.dst = &p.d4; o.bias = h->conv_b[12];
    if ((rc = r5_gemm(c, 12, p.h4, 16, 8, 8, 1, h->tc_conv[12], o, 12))) return rc;
    o = R5Out{}; o.dst = &p.z; o.bias = h->conv_b[13];
    if ((rc = r5_gemm(c, 13, p.d4, 3, 1, 2, 1, h->tc_conv[13], o, 13))) return rc;
  }
  // ---- encoder transformer ---------------------------------------------------------------------------------------------------
  const int T25 = maxlen[4];
  for (int l = 0; l < h->dbg_layers && T25 > 0; ++l) {
    const LayerDev& d = h->layer[l];
    dim3 lgrid((T25 + 7) / 8, B);
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z.off, ws + p.y.off, d.ln1_w, d.ln1_b, p.z.item_stride, dlen[4], T25, nullptr);
    h->launches++; mark(h, 14, st);
    R5Out o; o.dst = &p.qkv;
    if ((rc = r5_gemm(c, 0, p.y, 1, 1, 0, 0, h->tc_qkv[l], o, 15))) return rc;
    {
      const int ntiles = (T25 + kAttQT - 1) / kAttQT;
      const int nsplit = std::max(1, std::min(ntiles, (4 * h->num_sms + kHeads * B - 1) / (kHeads * B)));
      const int tps = (ntiles + nsplit - 1) / nsplit;
      dim3 agrid((ntiles + tps - 1) / tps, kHeads, B);
      swa_attention2_kernel<<<agrid, 256, kAtt2SmemBytes, st>>>(ws + p.qkv.off, p.qkv.item_stride, ws + p.att.off, p.att.item_stride,
                                                                h->rope_cos, h->rope_sin, dlen[4], T25, nullptr, tps);
      h->launches++; mark(h, 16, st);
      CUDA_TRY(h, cudaGetLastError());
    }
    o = R5Out{}; o.dst = &p.z; o.res = &p.z; o.scale = d.ls1;          // o_proj + LayerScale + residual, in place
    if ((rc = r5_gemm(c, 1, p.att, 1, 1, 0, 0, h->tc_o[l], o, 17))) return rc;
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z.off, ws + p.y.off, d.ln2_w, d.ln2_b, p.z.item_stride, dlen[4], T25, nullptr);
    h->launches++; mark(h, 14, st);
    o = R5Out{}; o.dst = &p.ffn; o.act = 1;                             // fc1 + GELU(erf)
    if ((rc = r5_gemm(c, 2, p.y, 1, 1, 0, 0, h->tc_fc1[l], o, 18))) return rc;
    o = R5Out{}; o.dst = &p.z; o.res = &p.z; o.scale = d.ls2;          // fc2 + LayerScale + residual, in place
    if ((rc = r5_gemm(c, 14, p.ffn, 1, 1, 0, 0, h->tc_fc2[l], o, 19))) return rc;
  }
  // ---- stride-2 downsample (replicate pad materialised as 3 extra rows), RVQ input projections ----------------------------
  const int T = maxlen[5];
  if (T25 > 0) {
    dim3 pgrid((T25 + 3 + 7) / 8, B);
    tc4::pad_replicate_kernel<<<pgrid, 256, 0, st>>>(ws + p.z.off, p.z.item_stride, ws + p.zp.off, p.zp.item_stride, dlen[4], T25);
    h->launches++; mark(h, 26, st);
    R5Out o; o.dst = &p.e;
    if ((rc = r5_gemm(c, 15, p.zp, 4, 2, 0, 0, h->tc_down, o, 20))) return rc;
  }
  if (d_latent_opt && T > 0) {
    const long long n = (long long)kHidden * p.rows[5];
    dim3 tgrid((unsigned)((n + 255) / 256), B);
    latent_transpose_kernel<<<tgrid, 256, 0, st>>>(ws + p.e.off, p.e.item_stride, d_latent_opt, p.rows[5], dlen[5], T);
    h->launches++; mark(h, 23, st);
  }
  if (T > 0) {
    R5Out o; o.dst = &p.rp;
    if ((rc = r5_gemm(c, 16, p.e, 1, 1, 0, 0, h->tc_proj, o, 21))) return rc;
    if (total_frames > 0) {
      rvqtc::Params q{};
      q.rproj = ws + p.rp.off; q.item_stride = p.rp.item_stride; q.embed = h->embed; q.enorm = h->enorm;
      q.codes = reinterpret_cast<long long*>(d_codes);
      q.K = K; q.T_out = p.rows[5]; q.len = dlen[5]; q.uniform_len = T; q.B = B; q.total_frames = total_frames;
      q.frame_prefix = dprefix;
      rvqtc::rvq_tc_kernel<<<(total_frames + rvqtc::kFrames - 1) / rvqtc::kFrames, rvqtc::kThreads, rvqtc::kSmem, st>>>(
          h->map_embed_hi, h->map_embed_lo, q);
      h->launches++; mark(h, 22, st);
    }
  }
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}
