"""ORACLE (test infrastructure, NOT product code): CPU/numpy restatement of the Mimi encode path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline leg may import this
module; the product package ``tokenize_audio_b200`` never does
(tests/test_host_logic.py::test_product_never_imports_the_oracle enforces it).

What it restates: ``transformers.MimiModel.encode`` and ``MimiModel.decode`` (transformers 5.5.0, a third-party dependency
of potsawee/tokenize-audio that is NOT vendored under /root/reference; call sites
REF/emilia-mimi/process_shard.py:81-84,124-127 and REF/*/utils.py:64). ``TF`` below means
``transformers/models/mimi/modeling_mimi.py``. Every function cites the lines it follows.

Parity pinning: the reference repository has no tests or golden vectors for this path
("parity unpinned by the reference"). This restatement is pinned instead
against outputs of the real ``transformers.MimiModel`` generated in the build container by
``tests/golden/make_golden.py`` (committed as ``tests/golden/*.npz``) and, wherever transformers
is importable, against a live ``MimiModel`` run on the same seeded weights
(tests/test_oracle_vs_transformers.py).

All arithmetic is float32 (the reference runs fp32 eager); only the final codes
are integers (int64).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np

try:  # exact erf for GELU; scipy is in the image. Fallback: math.erf loop (small cases only)
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float32])

F32 = np.float32

# (prefix, stride) of the 14 SEANet convs in execution order; TF:454-496 (MimiEncoder.__init__)
_SEANET = (
    ("encoder.layers.0", 1),
    ("encoder.layers.1.block.1", 1), ("encoder.layers.1.block.3", 1), ("encoder.layers.3", 4),
    ("encoder.layers.4.block.1", 1), ("encoder.layers.4.block.3", 1), ("encoder.layers.6", 5),
    ("encoder.layers.7.block.1", 1), ("encoder.layers.7.block.3", 1), ("encoder.layers.9", 6),
    ("encoder.layers.10.block.1", 1), ("encoder.layers.10.block.3", 1), ("encoder.layers.12", 8),
    ("encoder.layers.14", 1),
)
N_HEADS, HEAD_DIM, WINDOW, LN_EPS, ROPE_THETA = 8, 64, 250, 1e-5, 10000.0


def elu(x: np.ndarray) -> np.ndarray:
    """nn.ELU(alpha=1), TF:428,473,478: x if x > 0 else exp(x) - 1 (torch uses expm1-free form
    ``(exp(x) - 1)``; both agree to fp32 rounding)."""
    x = x.astype(F32, copy=False)
    return np.where(x > 0, x, np.expm1(np.minimum(x, F32(0)))).astype(F32)


def conv1d_causal(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int,
                  pad_mode: str = "constant", chunk: int = 1 << 15) -> np.ndarray:
    """MimiConv1d.forward, TF:331-351 with _get_extra_padding_for_conv1d TF:273-283 and _pad1d
    TF:285-301. ``x`` [C_in, L] -> [C_out, ceil(L/stride)].

    left pad = k - stride (TF:256 padding_total, causal branch TF:343); right pad
    = ceil(L/stride)*stride - L; zeros for pad_mode "constant", edge value for "replicate"
    (only the stride-2 downsample conv, TF:1422-1431)."""
    cout, cin, k = w.shape
    L = x.shape[1]
    n_out = -(-L // stride)
    pad_l = k - stride
    pad_r = n_out * stride - L          # == (n_frames*stride + k - padding_total) - L
    mode = "edge" if pad_mode == "replicate" else "constant"
    xp = np.pad(x.astype(F32, copy=False), ((0, 0), (pad_l, pad_r)), mode=mode)
    w2 = np.ascontiguousarray(w.reshape(cout, cin * k).astype(F32, copy=False))
    y = np.empty((cout, n_out), F32)
    sc, st = xp.strides
    for j0 in range(0, n_out, chunk):
        j1 = min(n_out, j0 + chunk)
        # cols[j, c, tau] = xp[c, j*stride + tau]
        cols = np.lib.stride_tricks.as_strided(
            xp[:, j0 * stride:], shape=(j1 - j0, cin, k), strides=(st * stride, sc, st), writeable=False)
        y[:, j0:j1] = w2 @ cols.reshape(j1 - j0, cin * k).T
    if b is not None:
        y += b.astype(F32)[:, None]
    return y


def seanet_encoder(sd: Dict[str, np.ndarray], x: np.ndarray, taps: Optional[dict] = None) -> np.ndarray:
    """MimiEncoder.forward TF:490-496 + MimiResnetBlock.forward TF:437-451. ``x`` [1, N] -> [512, T25].

    ELU is applied to the INPUT of the following conv; the resblock skip carries the
    un-activated tensor (identity shortcut, use_conv_shortcut=False)."""
    def conv(name, h, stride):
        return conv1d_causal(h, sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"], stride)

    h = conv("encoder.layers.0", x, 1)
    if taps is not None:
        taps["seanet.l0"] = h
    for blk, down, ratio in ((1, 3, 4), (4, 6, 5), (7, 9, 6), (10, 12, 8)):
        r = conv(f"encoder.layers.{blk}.block.1", elu(h), 1)
        r = conv(f"encoder.layers.{blk}.block.3", elu(r), 1)
        h = h + r
        if taps is not None:
            taps[f"seanet.res{blk}"] = h
        h = conv(f"encoder.layers.{down}", elu(h), ratio)
        if taps is not None:
            taps[f"seanet.down{down}"] = h
    h = conv("encoder.layers.14", elu(h), 1)
    if taps is not None:
        taps["seanet.out"] = h
    return h


def layer_norm(x: np.ndarray, w: np.ndarray, b: np.ndarray, eps: float = LN_EPS) -> np.ndarray:
    """nn.LayerNorm(512, eps=1e-5), TF:934-935 (biased variance)."""
    x = x.astype(F32, copy=False)
    mu = x.mean(axis=-1, keepdims=True, dtype=F32)
    xc = x - mu
    var = (xc * xc).mean(axis=-1, keepdims=True, dtype=F32)
    return (xc / np.sqrt(var + F32(eps)) * w.astype(F32) + b.astype(F32)).astype(F32)


def rope_tables(T: int) -> Tuple[np.ndarray, np.ndarray]:
    """MimiRotaryEmbedding TF:538-577: inv_freq = theta^(-2i/64) in fp32, positions arange(T),
    emb = cat(freqs, freqs); returns cos, sin of shape [T, 64] (fp32)."""
    inv_freq = (F32(1.0) / (F32(ROPE_THETA) ** (np.arange(0, HEAD_DIM, 2, dtype=np.int64).astype(F32) / F32(HEAD_DIM)))).astype(F32)
    freqs = np.arange(T, dtype=F32)[:, None] * inv_freq[None, :]
    emb = np.concatenate([freqs, freqs], axis=-1).astype(F32)
    return np.cos(emb).astype(F32), np.sin(emb).astype(F32)


def apply_rope(u: np.ndarray, cos: np.ndarray, sin: np.ndarray) -> np.ndarray:
    """apply_rotary_pos_emb + rotate_half TF:580-611: u*cos + [-u[32:], u[:32]]*sin. u [H,T,64]."""
    half = HEAD_DIM // 2
    rot = np.concatenate([-u[..., half:], u[..., :half]], axis=-1)
    return (u * cos[None] + rot * sin[None]).astype(F32)


def gelu_erf(x: np.ndarray) -> np.ndarray:
    """ACT2FN["gelu"] = exact erf GELU (config.hidden_act="gelu"), used by MimiMLP TF:614-627."""
    x = x.astype(F32, copy=False)
    return (F32(0.5) * x * (F32(1.0) + _erf(x * F32(1.0 / math.sqrt(2.0))).astype(F32))).astype(F32)


def sliding_window_attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, window: int = WINDOW) -> np.ndarray:
    """MimiSdpaAttention TF:852-916 with create_sliding_window_causal_mask TF:1096-1102
    (transformers/masking_utils.py:74-101): key j visible to query i iff j <= i and j > i - window;
    scale 1/sqrt(64); softmax in fp32. q,k,v [H,T,64] -> [H,T,64]."""
    H, T, D = q.shape
    i = np.arange(T)[:, None]
    j = np.arange(T)[None, :]
    allowed = (j <= i) & (j > i - window)
    out = np.empty_like(q, dtype=F32)
    scale = F32(1.0 / math.sqrt(D))
    for h in range(H):
        s = (q[h] @ k[h].T) * scale
        s = np.where(allowed, s, F32(-np.inf))
        s = s - s.max(axis=-1, keepdims=True)
        p = np.exp(s).astype(F32)
        p /= p.sum(axis=-1, keepdims=True, dtype=F32)
        out[h] = p @ v[h]
    return out


def encoder_transformer(sd: Dict[str, np.ndarray], z: np.ndarray, taps: Optional[dict] = None,
                        prefix: str = "encoder_transformer") -> np.ndarray:
    """MimiTransformerModel.forward TF:1015-1140 over MimiTransformerLayer.forward TF:939-993.
    ``z`` [T25, 512] -> [T25, 512]. Pre-LN blocks, LayerScale on both branches (TF:499-511),
    no linear biases, no final norm, positions arange(T25), no padding information."""
    T = z.shape[0]
    cos, sin = rope_tables(T)
    z = z.astype(F32, copy=True)
    for l in range(8):
        p = f"{prefix}.layers.{l}"
        y = layer_norm(z, sd[f"{p}.input_layernorm.weight"], sd[f"{p}.input_layernorm.bias"])
        def heads(w):
            return (y @ sd[w].T.astype(F32)).reshape(T, N_HEADS, HEAD_DIM).transpose(1, 0, 2)
        q = apply_rope(heads(f"{p}.self_attn.q_proj.weight"), cos, sin)
        k = apply_rope(heads(f"{p}.self_attn.k_proj.weight"), cos, sin)
        v = heads(f"{p}.self_attn.v_proj.weight")
        a = sliding_window_attention(q, k, v).transpose(1, 0, 2).reshape(T, N_HEADS * HEAD_DIM)
        o = a @ sd[f"{p}.self_attn.o_proj.weight"].T
        z = z + sd[f"{p}.self_attn_layer_scale.scale"] * o
        y = layer_norm(z, sd[f"{p}.post_attention_layernorm.weight"], sd[f"{p}.post_attention_layernorm.bias"])
        m = gelu_erf(y @ sd[f"{p}.mlp.fc1.weight"].T) @ sd[f"{p}.mlp.fc2.weight"].T
        z = (z + sd[f"{p}.mlp_layer_scale.scale"] * m).astype(F32)
        if taps is not None:
            taps[f"transformer.layer{l}"] = z
    return z


def codebook_embed(sd: Dict[str, np.ndarray], prefix: str) -> np.ndarray:
    """MimiEuclideanCodebook.embed TF:1191-1195: embed_sum / clamp(cluster_usage, min=1e-5)[:, None]."""
    usage = np.maximum(sd[f"{prefix}.codebook.cluster_usage"].astype(F32), F32(1e-5))
    return (sd[f"{prefix}.codebook.embed_sum"].astype(F32) / usage[:, None]).astype(F32)


def nearest_code(x: np.ndarray, embed: np.ndarray, return_margin: bool = False):
    """MimiEuclideanCodebook.quantize TF:1197-1202: cdist(p=2).argmin(-1). torch.cdist takes the
    matmul route (ATen _euclidean_dist): one GEMM of [-2x, |x|^2, 1] with [e, 1, |e|^2], then
    clamp_min(0).sqrt(); argmin returns the lowest index among equal minima. x [M,256]."""
    x = x.astype(F32, copy=False)
    xn = (x * x).sum(axis=1, keepdims=True, dtype=F32)
    en = (embed * embed).sum(axis=1, keepdims=True, dtype=F32)
    x_ = np.concatenate([F32(-2.0) * x, xn, np.ones_like(xn)], axis=1)
    y_ = np.concatenate([embed, np.ones_like(en), en], axis=1)
    d = np.sqrt(np.maximum(x_ @ y_.T, F32(0.0))).astype(F32)
    idx = d.argmin(axis=1).astype(np.int64)
    if not return_margin:
        return idx
    part = np.partition(d, 1, axis=1)[:, :2]
    margin = (part[:, 1] - part[:, 0]) / np.maximum(part[:, 0], F32(1e-30))
    return idx, margin.astype(F32)


def rvq_encode(sd: Dict[str, np.ndarray], e: np.ndarray, num_quantizers: int,
               margins: Optional[list] = None) -> np.ndarray:
    """MimiSplitResidualVectorQuantizer.encode TF:1311-1338 over
    MimiResidualVectorQuantizer.encode TF:1262-1280. ``e`` [512, T] -> codes [K, T] int64.

    One semantic stage on P_sem*e, then K-1 acoustic stages that start again from the
    UN-quantised e with their own input_proj (TF:1330-1336)."""
    if num_quantizers > 32:
        raise ValueError(
            "The number of quantizers (i.e codebooks) asked should be lower than the total number of "
            f"quantizers 32, but is currently {num_quantizers}.")
    if num_quantizers < 1:
        raise ValueError(
            "The number of quantizers (i.e codebooks) asked should be higher than the number of semantic "
            f"quantizers 1, but is currently {num_quantizers}.")
    codes: List[np.ndarray] = []
    for which, n in (("semantic", 1), ("acoustic", num_quantizers - 1)):
        p = f"quantizer.{which}_residual_vector_quantizer"
        proj = sd[f"{p}.input_proj.weight"][:, :, 0].astype(F32)          # [256, 512]
        r = (proj @ e.astype(F32)).T.copy()                               # [T, 256]
        for s in range(n):
            emb = codebook_embed(sd, f"{p}.layers.{s}")
            if margins is not None:
                idx, mg = nearest_code(r, emb, return_margin=True)
                margins.append(mg)
            else:
                idx = nearest_code(r, emb)
            r = (r - emb[idx]).astype(F32)                                # TF:1276-1277
            codes.append(idx)
    return np.stack(codes, axis=0)


def encoded_length(n_samples: int) -> int:
    """MimiModel.get_encoded_length TF:1490-1503: ceil through strides 4,5,6,8 then 2 == ceil(N/1920)."""
    L = int(n_samples)
    for s in (4, 5, 6, 8, 2):
        L = -(-L // s)
    return L


def encode(sd: Dict[str, np.ndarray], input_values: np.ndarray, num_quantizers: Optional[int] = None,
           padding_mask: Optional[np.ndarray] = None, taps: Optional[dict] = None,
           margins: Optional[list] = None) -> np.ndarray:
    """MimiModel.encode TF:1522-1611 / _encode_frame TF:1455-1488.
    ``input_values`` [B,1,N] fp32 -> audio_codes [B,K,T] int64, T = ceil(N/1920).
    ``padding_mask`` is accepted and ignored, exactly as the reference does (TF:1469,1472 TODOs)."""
    K = 32 if num_quantizers is None else int(num_quantizers)
    if K > 32:
        raise ValueError(
            "The number of quantizers (i.e codebooks) asked should be lower than the total number of "
            f"quantizers 32, but is currently {K}.")
    B, C, N = input_values.shape
    if C < 1 or C > 2:
        raise ValueError(f"Number of audio channels must be 1 or 2, but got {C}")
    out = []
    for b in range(B):
        t = {} if taps is not None else None
        h = seanet_encoder(sd, input_values[b].astype(F32), t)            # [512, T25]
        z = encoder_transformer(sd, h.T, t)                               # [T25, 512]
        e = conv1d_causal(z.T, sd["downsample.conv.weight"], None, 2, pad_mode="replicate")   # TF:1484
        if t is not None:
            t["latent"] = e
            for k_, v_ in t.items():
                taps.setdefault(k_, []).append(v_)
        out.append(rvq_encode(sd, e, K, margins))
    return np.stack(out, axis=0)


# ---- decode direction: MimiModel.decode TF:1613-1679 / _decode_frame TF:1594-1611 -----------------

def conv_transpose1d_causal(x: np.ndarray, w: np.ndarray, b: Optional[np.ndarray], stride: int, groups: int = 1) -> np.ndarray:
    """MimiConvTranspose1d.forward TF:403-409 with the causal trimming of TF:383-391 (trim_right_ratio = 1: the last
    k - stride outputs are dropped, none on the left). ``x`` [C_in, L], ``w`` [C_in, C_out/groups, k] (nn.ConvTranspose1d
    layout) -> [C_out, L * stride].  full[o, i*stride + tau] += x[c, i] * w[c, o, tau]."""
    cin, cog, k = w.shape
    L = x.shape[1]
    full_len = (L - 1) * stride + k
    x = x.astype(F32, copy=False)
    if groups == 1:
        full = np.zeros((cog, full_len), F32)
        for tau in range(k):
            full[:, tau: tau + (L - 1) * stride + 1: stride] += w[:, :, tau].T.astype(F32) @ x
    else:                                   # depthwise (groups == C_in == C_out): TF:1433-1441
        assert groups == cin and cog == 1
        full = np.zeros((cin, full_len), F32)
        for tau in range(k):
            full[:, tau: tau + (L - 1) * stride + 1: stride] += w[:, 0, tau:tau + 1].astype(F32) * x
    if b is not None:
        full += b.astype(F32)[:, None]
    return full[:, : full_len - (k - stride)]


def rvq_decode(sd: Dict[str, np.ndarray], codes: np.ndarray) -> np.ndarray:
    """MimiSplitResidualVectorQuantizer.decode TF:1340-1350 over MimiResidualVectorQuantizer.decode TF:1282-1297:
    codes [K, T] -> [512, T]: per RVQ the sum of the codebook rows, then that RVQ's output_proj (1x1 conv, no bias)."""
    out = np.zeros((512, codes.shape[1]), F32)
    for which, lo, hi in (("semantic", 0, 1), ("acoustic", 1, codes.shape[0])):
        if hi <= lo:
            continue
        p = f"quantizer.{which}_residual_vector_quantizer"
        q = np.zeros((codes.shape[1], 256), F32)
        for s in range(lo, hi):
            q = q + codebook_embed(sd, f"{p}.layers.{s - lo}")[codes[s]]
        out = out + (sd[f"{p}.output_proj.weight"][:, :, 0].astype(F32) @ q.T)
    return out.astype(F32)


def seanet_decoder(sd: Dict[str, np.ndarray], z: np.ndarray) -> np.ndarray:
    """MimiDecoder.forward TF:1169-1174 (layers built at TF:1146-1167): conv k7, then per ratio (8, 6, 5, 4)
    ELU -> ConvTranspose1d(k = 2r, stride r) -> residual block, then ELU -> conv k3 to one channel. [512, T25] -> [1, 960*T25]."""
    def conv(name, h, stride=1):
        return conv1d_causal(h, sd[f"{name}.conv.weight"], sd[f"{name}.conv.bias"], stride)
    h = conv("decoder.layers.0", z)
    for up, res, r in (("decoder.layers.2", "decoder.layers.3", 8), ("decoder.layers.5", "decoder.layers.6", 6),
                       ("decoder.layers.8", "decoder.layers.9", 5), ("decoder.layers.11", "decoder.layers.12", 4)):
        h = conv_transpose1d_causal(elu(h), sd[f"{up}.conv.weight"], sd[f"{up}.conv.bias"], r)
        t = conv(f"{res}.block.1", elu(h))
        t = conv(f"{res}.block.3", elu(t))
        h = h + t
    return conv("decoder.layers.14", elu(h))


def decode(sd: Dict[str, np.ndarray], audio_codes: np.ndarray) -> np.ndarray:
    """MimiModel.decode: ``audio_codes`` [B, K, T] int -> audio_values [B, 1, 1920*T] fp32."""
    out = []
    for b in range(audio_codes.shape[0]):
        e = rvq_decode(sd, np.asarray(audio_codes[b]))
        u = conv_transpose1d_causal(e, sd["upsample.conv.weight"], None, 2, groups=512)          # TF:1596
        z = encoder_transformer(sd, u.T, prefix="decoder_transformer").T                           # TF:1597-1604
        out.append(seanet_decoder(sd, z))
    return np.stack(out, axis=0)
