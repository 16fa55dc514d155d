// Host side of the decode direction (included by mimi_b200.cu): weight packing and the launch sequence of MimiModel.decode
// (modeling_mimi.py:1613-1679 -> _decode_frame :1594-1611). Everything runs on the exact-fp32 FFMA kernels.
//
// A causal ConvTranspose1d(C_in -> C_out, k = 2r, stride r) whose last k - r outputs are trimmed (MimiConvTranspose1d,
// modeling_mimi.py:354-409, trim_right_ratio = 1) is, in channels-last layout, a causal conv with TWO taps and r * C_out
// output columns:   y[i*r + phi, o] = b[o] + sum_c x[i, c] * w[c, o, phi] + x[i-1, c] * w[c, o, phi + r]
// i.e. out[i, phi*C_out + o] over the rows (i-1, i): the [T_in][r * C_out] result IS the channels-last [T_in * r][C_out] tensor.

namespace {
struct DecConvT { int cin, cout, r; };
const DecConvT kDecUp[4] = {{1024, 512, 8}, {512, 256, 6}, {256, 128, 5}, {128, 64, 4}};

struct PlanDec {
  long long q, e, z, y, qkv, att, ffn, c0, t[4], r[4], o32;      // float offsets
  size_t bytes = 0;
};

PlanDec make_plan_dec(int B, long long T) {
  PlanDec p;
  long long off = 0;
  auto take = [&](long long n) { const long long o = off; off += (n + 63) / 64 * 64; return o; };
  const long long T25 = 2 * T;
  p.q = take((long long)B * T * 512);    p.e = take((long long)B * T * 512);
  p.z = take((long long)B * T25 * 512);  p.y = take((long long)B * T25 * 512);
  p.qkv = take((long long)B * T25 * 1536); p.att = take((long long)B * T25 * 512); p.ffn = take((long long)B * T25 * 2048);
  p.c0 = take((long long)B * T25 * 1024);
  long long rows = T25;
  for (int s = 0; s < 4; ++s) {
    rows *= kDecUp[s].r;
    p.t[s] = take((long long)B * rows * kDecUp[s].cout);
    p.r[s] = take((long long)B * rows * (kDecUp[s].cout / 2));
  }
  p.o32 = take((long long)B * rows * 32);
  p.bytes = (size_t)off * sizeof(float) + 256;
  return p;
}
}  // namespace

static int dec_load_weights(mimi_b200* h, const mimi_b200_decoder_weights_t* w) {
  int rc;
  auto vec = [](const float* p, size_t n) { return std::vector<float>(p, p + n); };
  for (void* p : h->dec_allocs) cudaFree(p);
  h->dec_allocs.clear();
  h->dec_loaded = false;
  const size_t mark = h->allocs.size();          // dev_upload() records into h->allocs: move the new entries afterwards
  auto need = [&](const void* p) { return p != nullptr; };
  if (!need(w->semantic_output_proj_weight) || !need(w->acoustic_output_proj_weight) || !need(w->upsample_weight) ||
      !need(w->conv_in_weight) || !need(w->conv_in_bias) || !need(w->conv_out_weight) || !need(w->conv_out_bias))
    return fail(h, MIMI_B200_ERR_ARG, "load_decoder_weights: missing tensor");
  {
    // both output projections as one GEMM over q = [sem (256) | aco (256)]: Wt[k][n], k < 256 semantic, k >= 256 acoustic
    std::vector<float> wt((size_t)512 * 512);
    for (int n = 0; n < 512; ++n)
      for (int k = 0; k < 256; ++k) {
        wt[(size_t)k * 512 + n] = w->semantic_output_proj_weight[(size_t)n * 256 + k];
        wt[(size_t)(256 + k) * 512 + n] = w->acoustic_output_proj_weight[(size_t)n * 256 + k];
      }
    if ((rc = dev_upload(h, &h->dec_proj_wt, wt))) return rc;
  }
  if ((rc = dev_upload(h, &h->dec_up_w, vec(w->upsample_weight, 512 * 4)))) return rc;
  for (int l = 0; l < MIMI_B200_NUM_LAYERS; ++l) {
    const mimi_b200_layer_weights_t& s = w->layer[l];
    const float* req[] = {s.input_layernorm_weight, s.input_layernorm_bias, s.q_proj_weight, s.k_proj_weight, s.v_proj_weight,
                          s.o_proj_weight, s.self_attn_layer_scale, s.post_attention_layernorm_weight,
                          s.post_attention_layernorm_bias, s.fc1_weight, s.fc2_weight, s.mlp_layer_scale};
    for (const float* q : req) if (!q) return fail(h, MIMI_B200_ERR_ARG, "load_decoder_weights: missing transformer tensor");
    LayerDev& d = h->dlayer[l];
    if ((rc = dev_upload(h, &d.ln1_w, vec(s.input_layernorm_weight, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ln1_b, vec(s.input_layernorm_bias, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ln2_w, vec(s.post_attention_layernorm_weight, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ln2_b, vec(s.post_attention_layernorm_bias, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ls1, vec(s.self_attn_layer_scale, 512)))) return rc;
    if ((rc = dev_upload(h, &d.ls2, vec(s.mlp_layer_scale, 512)))) return rc;
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
  if ((rc = dev_upload(h, &h->dec_in_wt, pack_conv(w->conv_in_weight, 1024, 512, 7)))) return rc;
  if ((rc = dev_upload(h, &h->dec_in_b, vec(w->conv_in_bias, 1024)))) return rc;
  for (int s = 0; s < 4; ++s) {
    const DecConvT& g = kDecUp[s];
    if (!w->up_weight[s] || !w->up_bias[s] || !w->res_a_weight[s] || !w->res_a_bias[s] || !w->res_b_weight[s] || !w->res_b_bias[s])
      return fail(h, MIMI_B200_ERR_ARG, "load_decoder_weights: missing SEANet decoder tensor");
    // ConvTranspose1d weight [C_in][C_out][2r] -> Wt[(tap*C_in + c)][(phi*C_out + o)]: tap 0 = row i-1 (w[.., phi + r]), tap 1 = row i
    const int N = g.r * g.cout;
    std::vector<float> wt((size_t)2 * g.cin * N), bias((size_t)N);
    for (int c = 0; c < g.cin; ++c)
      for (int o = 0; o < g.cout; ++o)
        for (int phi = 0; phi < g.r; ++phi) {
          const float* wp = w->up_weight[s] + ((size_t)c * g.cout + o) * (2 * g.r);
          wt[((size_t)0 * g.cin + c) * N + phi * g.cout + o] = wp[phi + g.r];
          wt[((size_t)1 * g.cin + c) * N + phi * g.cout + o] = wp[phi];
        }
    for (int phi = 0; phi < g.r; ++phi)
      for (int o = 0; o < g.cout; ++o) bias[(size_t)phi * g.cout + o] = w->up_bias[s][o];
    if ((rc = dev_upload(h, &h->dec_up_wt[s], wt))) return rc;
    if ((rc = dev_upload(h, &h->dec_up_b[s], bias))) return rc;
    if ((rc = dev_upload(h, &h->dec_ra_wt[s], pack_conv(w->res_a_weight[s], g.cout / 2, g.cout, 3)))) return rc;
    if ((rc = dev_upload(h, &h->dec_ra_b[s], vec(w->res_a_bias[s], g.cout / 2)))) return rc;
    if ((rc = dev_upload(h, &h->dec_rb_wt[s], pack_conv(w->res_b_weight[s], g.cout, g.cout / 2, 1)))) return rc;
    if ((rc = dev_upload(h, &h->dec_rb_b[s], vec(w->res_b_bias[s], g.cout)))) return rc;
  }
  {
    // last conv 64 -> 1, k3, as a 32-column GEMM whose columns 1..31 are zero
    std::vector<float> wt((size_t)3 * 64 * 32, 0.f), bias(32, 0.f);
    for (int ci = 0; ci < 64; ++ci)
      for (int tau = 0; tau < 3; ++tau) wt[((size_t)tau * 64 + ci) * 32] = w->conv_out_weight[(size_t)ci * 3 + tau];
    bias[0] = w->conv_out_bias[0];
    if ((rc = dev_upload(h, &h->dec_out_wt, wt))) return rc;
    if ((rc = dev_upload(h, &h->dec_out_b, bias))) return rc;
  }
  if (!h->dec_bad) CUDA_TRY(h, cudaMalloc((void**)&h->dec_bad, sizeof(int)));
  h->dec_allocs.assign(h->allocs.begin() + mark, h->allocs.end());
  h->allocs.resize(mark);
  h->dec_loaded = true;
  return MIMI_B200_OK;
}

static int decode_impl(mimi_b200* h, const int64_t* d_codes, int B, int K, int64_t T, float* d_audio, void* d_workspace,
                       size_t workspace_bytes, cudaStream_t st) {
  const PlanDec p = make_plan_dec(B, T);
  uintptr_t base = (reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255);
  if (base + p.bytes - 256 > reinterpret_cast<uintptr_t>(d_workspace) + workspace_bytes)
    return fail(h, MIMI_B200_ERR_WORKSPACE, "decode: workspace too small, need " + std::to_string(p.bytes));
  float* ws = reinterpret_cast<float*>(base);
  int rc;
  const int T12 = (int)T, T25 = (int)(2 * T);
  CUDA_TRY(h, cudaMemsetAsync(h->dec_bad, 0, sizeof(int), st));
  // ---- quantizer.decode: codebook lookups summed per RVQ, then both output projections as one GEMM ------------------
  rvq_decode_sum_kernel<<<(unsigned)(((long long)B * T + 7) / 8), 256, 0, st>>>(reinterpret_cast<const long long*>(d_codes), B, K, T,
                                                                               h->embed, ws + p.q, h->dec_bad);
  h->launches++;
  auto gemm = [&](const float* A, long long a_stride, int cin, int k, int pad, const float* wt, const float* bias, int N,
                  float* out, long long out_stride, int rows, int elu_in, const float* res = nullptr, const float* scale = nullptr,
                  int act = 0) {
    GemmParams g{};
    g.A = A; g.Wt = wt; g.bias = bias; g.scale = scale; g.res = res; g.out = out; g.len_in = nullptr; g.uniform_len_in = rows;
    g.a_item_stride = a_stride; g.out_item_stride = out_stride; g.Cin = cin; g.stride = 1; g.pad_left = pad; g.K = k * cin; g.N = N;
    g.elu_in = elu_in; g.act = act;
    return launch_gemm(h, g, B, rows, st, -1);
  };
  if ((rc = gemm(ws + p.q, (long long)T12 * 512, 512, 1, 0, h->dec_proj_wt, nullptr, 512, ws + p.e, (long long)T12 * 512, T12, 0))) return rc;
  // ---- upsample 12.5 -> 25 Hz ------------------------------------------------------------------------------------------
  {
    const long long n = (long long)B * T25 * 512;
    upsample2_depthwise_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws + p.e, T, h->dec_up_w, ws + p.z, n);
    h->launches++;
  }
  // ---- decoder transformer (same block as the encoder's, modeling_mimi.py:926-1140) ------------------------------------------
  const long long s512 = (long long)T25 * 512;
  for (int l = 0; l < MIMI_B200_NUM_LAYERS; ++l) {
    const LayerDev& d = h->dlayer[l];
    dim3 lgrid((T25 + 7) / 8, B);
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z, ws + p.y, d.ln1_w, d.ln1_b, s512, nullptr, T25, nullptr);
    h->launches++;
    if ((rc = gemm(ws + p.y, s512, 512, 1, 0, d.qkv_wt, nullptr, 1536, ws + p.qkv, (long long)T25 * 1536, T25, 0))) return rc;
    dim3 agrid((T25 + kAttQT - 1) / kAttQT, kHeads, B);
    swa_attention_kernel<<<agrid, 256, kAttSmemBytes, st>>>(ws + p.qkv, (long long)T25 * 1536, ws + p.att, s512, h->rope_cos, h->rope_sin,
                                                            nullptr, T25);
    h->launches++;
    CUDA_TRY(h, cudaGetLastError());
    if ((rc = gemm(ws + p.att, s512, 512, 1, 0, d.o_wt, nullptr, 512, ws + p.z, s512, T25, 0, ws + p.z, d.ls1))) return rc;
    layernorm512_kernel<<<lgrid, 256, 0, st>>>(ws + p.z, ws + p.y, d.ln2_w, d.ln2_b, s512, nullptr, T25, nullptr);
    h->launches++;
    if ((rc = gemm(ws + p.y, s512, 512, 1, 0, d.fc1_wt, nullptr, 2048, ws + p.ffn, (long long)T25 * 2048, T25, 0, nullptr, nullptr, 1))) return rc;
    if ((rc = gemm(ws + p.ffn, (long long)T25 * 2048, 2048, 1, 0, d.fc2_wt, nullptr, 512, ws + p.z, s512, T25, 0, ws + p.z, d.ls2))) return rc;
  }
  // ---- SEANet decoder (MimiDecoder, modeling_mimi.py:1143-1174) ----------------------------------------------------------
  if ((rc = gemm(ws + p.z, s512, 512, 7, 6, h->dec_in_wt, h->dec_in_b, 1024, ws + p.c0, (long long)T25 * 1024, T25, 0))) return rc;
  const float* cur = ws + p.c0;
  long long rows = T25;
  int C = 1024;
  for (int s = 0; s < 4; ++s) {
    const DecConvT& g = kDecUp[s];
    // ELU -> ConvTranspose1d(C -> C/2, k = 2r, stride r) as a 2-tap causal conv with r * C/2 columns
    if ((rc = gemm(cur, rows * C, C, 2, 1, h->dec_up_wt[s], h->dec_up_b[s], g.r * g.cout, ws + p.t[s], rows * g.r * g.cout, (int)rows, 1))) return rc;
    rows *= g.r;
    C = g.cout;
    // residual block: ELU -> conv k3 (C -> C/2) -> ELU -> conv k1 (C/2 -> C) + skip (modeling_mimi.py:412-451)
    if ((rc = gemm(ws + p.t[s], rows * C, C, 3, 2, h->dec_ra_wt[s], h->dec_ra_b[s], C / 2, ws + p.r[s], rows * (C / 2), (int)rows, 1))) return rc;
    if ((rc = gemm(ws + p.r[s], rows * (C / 2), C / 2, 1, 0, h->dec_rb_wt[s], h->dec_rb_b[s], C, ws + p.t[s], rows * C, (int)rows, 1,
                   ws + p.t[s]))) return rc;
    cur = ws + p.t[s];
  }
  // ELU -> conv 64 -> 1, k3
  if ((rc = gemm(cur, rows * 64, 64, 3, 2, h->dec_out_wt, h->dec_out_b, 32, ws + p.o32, rows * 32, (int)rows, 1))) return rc;
  {
    const long long n = (long long)B * rows;
    take_column0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ws + p.o32, d_audio, n);
    h->launches++;
  }
  CUDA_TRY(h, cudaGetLastError());
  return MIMI_B200_OK;
}
