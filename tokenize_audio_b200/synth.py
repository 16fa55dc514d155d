"""Deterministic synthetic inputs for the Mimi encode path: weights and speech-shaped audio.

Real ``kyutai/mimi`` weights are not available offline, so parity tests, ``smoke()`` and
``bench.py`` all run on a seeded synthetic state dict that has the exact key names, shapes and
dtypes of ``transformers.MimiModel.state_dict()`` (encode side only).
Everything here is a pure function of the seed (numpy PCG64), so the same tensors are
regenerated bit-for-bit in this container and on the GPU box.

The generator deliberately does NOT run the model to derive anything (codebooks are drawn from
the RNG with hard-coded scales), so no floating-point summation order can leak into the weights.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np

# ---- architecture constants (transformers MimiConfig() defaults == kyutai/mimi;
#      TF/models/mimi/configuration_mimi.py:100-184) -------------------------------------------
SAMPLE_RATE = 24000
NUM_FILTERS = 64
RATIOS = (4, 5, 6, 8)                 # encoder order (reversed upsampling_ratios [8,6,5,4])
HIDDEN = 512
N_LAYERS = 8
N_HEADS = 8
HEAD_DIM = 64
FFN = 2048
SLIDING_WINDOW = 250
CODEBOOK_SIZE = 2048
CODEBOOK_DIM = 256
NUM_QUANTIZERS = 32
NUM_SEMANTIC = 1
FRAME_SIZE = 1920                     # 24 kHz samples per 12.5 Hz code frame

# (state-dict prefix, C_in, C_out, kernel, stride) for the 14 SEANet convs, in execution order
SEANET_CONVS = (
    ("encoder.layers.0", 1, 64, 7, 1),
    ("encoder.layers.1.block.1", 64, 32, 3, 1),
    ("encoder.layers.1.block.3", 32, 64, 1, 1),
    ("encoder.layers.3", 64, 128, 8, 4),
    ("encoder.layers.4.block.1", 128, 64, 3, 1),
    ("encoder.layers.4.block.3", 64, 128, 1, 1),
    ("encoder.layers.6", 128, 256, 10, 5),
    ("encoder.layers.7.block.1", 256, 128, 3, 1),
    ("encoder.layers.7.block.3", 128, 256, 1, 1),
    ("encoder.layers.9", 256, 512, 12, 6),
    ("encoder.layers.10.block.1", 512, 256, 3, 1),
    ("encoder.layers.10.block.3", 256, 512, 1, 1),
    ("encoder.layers.12", 512, 1024, 16, 8),
    ("encoder.layers.14", 1024, 512, 3, 1),
)


def _rng(seed: int, *stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([int(seed), *[int(s) for s in stream]]))


def _randn(g: np.random.Generator, *shape: int) -> np.ndarray:
    return g.standard_normal(size=shape, dtype=np.float32)


def synth_state_dict(seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded encode-side Mimi state dict as ``{name: np.float32 array}``.

    Scales are chosen so that activations stay O(1) through the 14 convs (variance-preserving
    gains for ELU inputs), LayerScale is large enough for the transformer to matter, and the
    codebooks have the same per-dimension scale as the residuals they quantise, so that argmin
    margins (and therefore near-tie flips) are realistic rather than degenerate. A handful of
    ``cluster_usage`` entries sit below the 1e-5 clamp of ``MimiEuclideanCodebook.embed``
    (TF/models/mimi/modeling_mimi.py:1191-1195) to exercise it.
    """
    sd: Dict[str, np.ndarray] = {}
    for li, (name, cin, cout, k, _s) in enumerate(SEANET_CONVS):
        g = _rng(seed, 1, li)
        fan_in = cin * k
        gain = 1.0 if li == 0 else 1.45          # ELU roughly halves the second moment
        if name.endswith("block.3"):
            gain = 0.8                           # residual branch: keep h + r from blowing up
        sd[f"{name}.conv.weight"] = _randn(g, cout, cin, k) * np.float32(gain / math.sqrt(fan_in))
        sd[f"{name}.conv.bias"] = _randn(g, cout) * np.float32(0.05)
    # the first conv sees raw audio with rms ~0.05-0.2: lift it to O(1)
    sd["encoder.layers.0.conv.weight"] *= np.float32(8.0)

    for l in range(N_LAYERS):
        g = _rng(seed, 2, l)
        p = f"encoder_transformer.layers.{l}"
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[f"{p}.self_attn.{nm}.weight"] = _randn(g, HIDDEN, HIDDEN) * np.float32(1.0 / math.sqrt(HIDDEN))
        # sharper attention than 1/sqrt(d) random init gives: scale q so logits have std ~2
        sd[f"{p}.self_attn.q_proj.weight"] *= np.float32(2.0)
        sd[f"{p}.mlp.fc1.weight"] = _randn(g, FFN, HIDDEN) * np.float32(1.0 / math.sqrt(HIDDEN))
        sd[f"{p}.mlp.fc2.weight"] = _randn(g, HIDDEN, FFN) * np.float32(1.0 / math.sqrt(FFN))
        for nm in ("input_layernorm", "post_attention_layernorm"):
            sd[f"{p}.{nm}.weight"] = np.float32(1.0) + _randn(g, HIDDEN) * np.float32(0.1)
            sd[f"{p}.{nm}.bias"] = _randn(g, HIDDEN) * np.float32(0.1)
        for nm in ("self_attn_layer_scale", "mlp_layer_scale"):
            sd[f"{p}.{nm}.scale"] = g.uniform(0.05, 0.5, size=HIDDEN).astype(np.float32)

    g = _rng(seed, 3)
    sd["downsample.conv.weight"] = _randn(g, HIDDEN, HIDDEN, 4) * np.float32(1.0 / math.sqrt(HIDDEN * 4))

    def codebook(prefix: str, g: np.random.Generator, sigma: float) -> None:
        usage = g.uniform(0.5, 50.0, size=CODEBOOK_SIZE).astype(np.float32)
        dead = g.choice(CODEBOOK_SIZE, size=4, replace=False)
        usage[dead] = np.float32(0.0)            # below the 1e-5 clamp -> |embed| ~1e5*|embed_sum|
        embed = _randn(g, CODEBOOK_SIZE, CODEBOOK_DIM) * np.float32(sigma)
        esum = embed * usage[:, None]
        esum[dead] = _randn(g, 4, CODEBOOK_DIM) * np.float32(sigma)   # -> far-away centroids
        sd[f"{prefix}.codebook.embed_sum"] = esum.astype(np.float32)
        sd[f"{prefix}.codebook.cluster_usage"] = usage
        sd[f"{prefix}.codebook.initialized"] = np.ones(1, np.float32)

    for which, n_layers, stream in (("semantic", NUM_SEMANTIC, 4), ("acoustic", NUM_QUANTIZERS - NUM_SEMANTIC, 5)):
        p = f"quantizer.{which}_residual_vector_quantizer"
        g = _rng(seed, stream)
        sd[f"{p}.input_proj.weight"] = _randn(g, CODEBOOK_DIM, HIDDEN, 1) * np.float32(1.0 / math.sqrt(HIDDEN))
        for s in range(n_layers):
            # residual rms per dimension shrinks slowly with random codebooks in 256-d
            codebook(f"{p}.layers.{s}", _rng(seed, stream, s), RVQ_SIGMA0 * (RVQ_DECAY ** s))
    return sd


# per-dimension rms of the projected latent (measured once on the synthetic model, seed 0, see
# tests/golden/make_golden.py --stats) and the per-stage shrink factor of the RVQ residual
RVQ_SIGMA0 = 1.5
RVQ_DECAY = 0.985


def state_dict_digest(sd: Dict[str, np.ndarray]) -> str:
    """sha256 over names + bytes, used by the golden fixtures to detect generator drift."""
    import hashlib

    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


def synth_speech(seed: int, n_samples: int, sr: int = SAMPLE_RATE) -> np.ndarray:
    """Speech-shaped fp32 mono signal: harmonic glottal-like source with
    an f0 random walk in 80-250 Hz, 3-4 slowly moving formant resonances, syllabic amplitude
    modulation at 3-6 Hz, unvoiced noise bursts, pauses with a -70 dBFS floor, peak level drawn
    uniformly in -20..-3 dBFS."""
    g = _rng(seed, 77)
    n = int(n_samples)
    if n <= 0:
        return np.zeros(0, np.float32)
    t = np.arange(n, dtype=np.float64) / sr
    # control-rate tracks at 100 Hz, linearly interpolated
    nc = n // (sr // 100) + 2
    tc = np.arange(nc, dtype=np.float64) / 100.0

    def track(lo, hi, step):
        x = np.cumsum(g.standard_normal(nc) * step)
        x = (x - x.min()) / max(np.ptp(x), 1e-9)
        return np.interp(t, tc, lo + (hi - lo) * x)

    f0 = track(80.0, 250.0, 1.0)
    phase = 2.0 * np.pi * np.cumsum(f0) / sr
    src = np.zeros(n)
    n_harm = 40
    formants = [track(c * 0.7, c * 1.3, 1.0) for c in (500.0, 1500.0, 2500.0, 3500.0)[: int(g.integers(3, 5))]]
    bw = 120.0
    for h in range(1, n_harm + 1):
        fh = f0 * h
        gain = np.zeros(n)
        for fc in formants:
            gain += 1.0 / (1.0 + ((fh - fc) / bw) ** 2)
        gain *= (fh < 0.45 * sr) / h
        src += gain * np.sin(h * phase + g.uniform(0, 2 * np.pi))
    am_f = g.uniform(3.0, 6.0)
    am = 0.55 + 0.45 * np.sin(2 * np.pi * am_f * t + g.uniform(0, 2 * np.pi))
    voiced = src * am
    # unvoiced bursts: differenced (high-passed) white noise, 10-20 % of the time
    noise = np.diff(g.standard_normal(n + 1))
    # segment decisions at 8 Hz (125 ms granules), smoothed by the interpolation
    ns = n // (sr // 8) + 2
    ts = np.arange(ns, dtype=np.float64) / 8.0
    seg = np.interp(t, ts, (g.uniform(size=ns) < g.uniform(0.10, 0.20)).astype(np.float64))
    pause = np.interp(t, ts, (g.uniform(size=ns) < g.uniform(0.05, 0.15)).astype(np.float64))
    x = voiced * (1.0 - seg) + 0.3 * np.std(voiced + 1e-12) * noise * seg
    x = x * (1.0 - pause)
    peak = np.max(np.abs(x)) + 1e-12
    level = 10.0 ** (g.uniform(-20.0, -3.0) / 20.0)
    x = x / peak * level + 10.0 ** (-70.0 / 20.0) * g.standard_normal(n) * (pause > 0.5)
    return x.astype(np.float32)


# (state-dict prefix, C_in, C_out, ratio) of the four ConvTranspose1d stages of the SEANet decoder and the index of the
# residual block that follows each (MimiDecoder.__init__, TF/models/mimi/modeling_mimi.py:1143-1167)
SEANET_DECODER_UPS = (
    ("decoder.layers.2", 1024, 512, 8, "decoder.layers.3"),
    ("decoder.layers.5", 512, 256, 6, "decoder.layers.6"),
    ("decoder.layers.8", 256, 128, 5, "decoder.layers.9"),
    ("decoder.layers.11", 128, 64, 4, "decoder.layers.12"),
)


def decoder_state_dict(seed: int = 0) -> Dict[str, np.ndarray]:
    """Seeded DECODE-side tensors of ``MimiModel.state_dict()`` (output projections, upsample, decoder transformer, SEANet
    decoder) for the decode direction. Kept apart from :func:`synth_state_dict` so that the
    encode-side digest the golden fixtures were made with never moves; merge the two dicts to get a full model."""
    sd: Dict[str, np.ndarray] = {}
    g = _rng(seed, 21)
    for which in ("semantic", "acoustic"):
        sd[f"quantizer.{which}_residual_vector_quantizer.output_proj.weight"] = \
            _randn(g, HIDDEN, CODEBOOK_DIM, 1) * np.float32(1.0 / math.sqrt(CODEBOOK_DIM))
    sd["upsample.conv.weight"] = (np.float32(0.5) + _randn(g, HIDDEN, 1, 4) * np.float32(0.3)).astype(np.float32)
    for l in range(N_LAYERS):
        g = _rng(seed, 22, l)
        p = f"decoder_transformer.layers.{l}"
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[f"{p}.self_attn.{nm}.weight"] = _randn(g, HIDDEN, HIDDEN) * np.float32(1.0 / math.sqrt(HIDDEN))
        sd[f"{p}.self_attn.q_proj.weight"] *= np.float32(2.0)
        sd[f"{p}.mlp.fc1.weight"] = _randn(g, FFN, HIDDEN) * np.float32(1.0 / math.sqrt(HIDDEN))
        sd[f"{p}.mlp.fc2.weight"] = _randn(g, HIDDEN, FFN) * np.float32(1.0 / math.sqrt(FFN))
        for nm in ("input_layernorm", "post_attention_layernorm"):
            sd[f"{p}.{nm}.weight"] = np.float32(1.0) + _randn(g, HIDDEN) * np.float32(0.1)
            sd[f"{p}.{nm}.bias"] = _randn(g, HIDDEN) * np.float32(0.1)
        for nm in ("self_attn_layer_scale", "mlp_layer_scale"):
            sd[f"{p}.{nm}.scale"] = g.uniform(0.05, 0.5, size=HIDDEN).astype(np.float32)
    g = _rng(seed, 23)
    sd["decoder.layers.0.conv.weight"] = _randn(g, 1024, HIDDEN, 7) * np.float32(0.5 / math.sqrt(HIDDEN * 7))
    sd["decoder.layers.0.conv.bias"] = _randn(g, 1024) * np.float32(0.05)
    for si, (name, cin, cout, r, res) in enumerate(SEANET_DECODER_UPS):
        g = _rng(seed, 24, si)
        # ConvTranspose1d weight layout [C_in, C_out, k]; every output sample sums two taps of every input channel
        sd[f"{name}.conv.weight"] = _randn(g, cin, cout, 2 * r) * np.float32(1.45 / math.sqrt(2 * cin))
        sd[f"{name}.conv.bias"] = _randn(g, cout) * np.float32(0.05)
        sd[f"{res}.block.1.conv.weight"] = _randn(g, cout // 2, cout, 3) * np.float32(1.45 / math.sqrt(cout * 3))
        sd[f"{res}.block.1.conv.bias"] = _randn(g, cout // 2) * np.float32(0.05)
        sd[f"{res}.block.3.conv.weight"] = _randn(g, cout, cout // 2, 1) * np.float32(0.8 / math.sqrt(cout // 2))
        sd[f"{res}.block.3.conv.bias"] = _randn(g, cout) * np.float32(0.05)
    g = _rng(seed, 25)
    sd["decoder.layers.14.conv.weight"] = _randn(g, 1, 64, 3) * np.float32(0.15 / math.sqrt(64 * 3))
    sd["decoder.layers.14.conv.bias"] = _randn(g, 1) * np.float32(0.01)
    return sd


def variant_state_dict(kind: str) -> Dict[str, np.ndarray]:
    """Adversarial weight draws for the parity tests (same key names / shapes as :func:`synth_state_dict`).

    ``"ties"``  seed-0 weights whose codebooks are full of EXACT ties: in every codebook rows 1024..2047 are bitwise
                copies of rows 0..1023 (embed_sum and cluster_usage both), so every winner has an identical twin at
                index + 1024 and ``argmin`` must return the lower one (TF/models/mimi/modeling_mimi.py:1200-1201,
                lowest index among equal minima); acoustic stages 3, 7 and 20..31 additionally hold all-zero rows
                (dead entries, embed == 0) at indices 5, 700, 1029, 1724 in a codebook scaled x3, where the zero
                vector is the nearest centroid for most frames and index 5 must win.
    ``"heavy"`` seed-1 weights with heavy tails: Student-t (3 degrees of freedom) conv / linear weights at the variance
                of the Gaussian draw, LayerScale 0.01 (the kyutai/mimi initial value), three x50 outlier output channels
                in each strided conv D1..D3 and four x50 outlier rows in every fc1 -- wide dynamic range inside one
                GEMM row, which is where a split-precision scheme loses bits first.
    """
    if kind == "ties":
        sd = synth_state_dict(0)
        half = CODEBOOK_SIZE // 2
        for which, n_layers in (("semantic", NUM_SEMANTIC), ("acoustic", NUM_QUANTIZERS - NUM_SEMANTIC)):
            for s in range(n_layers):
                p = f"quantizer.{which}_residual_vector_quantizer.layers.{s}.codebook"
                es, us = sd[f"{p}.embed_sum"].copy(), sd[f"{p}.cluster_usage"].copy()
                if which == "acoustic" and (s in (3, 7) or s >= 20):
                    es *= np.float32(3.0)
                    es[[5, 700]] = np.float32(0.0)
                    us[[5, 700]] = np.float32(1.0)
                es[half:] = es[:half]
                us[half:] = us[:half]
                sd[f"{p}.embed_sum"], sd[f"{p}.cluster_usage"] = es, us
        return sd
    if kind == "heavy":
        seed = 1
        sd = synth_state_dict(seed)

        def student(g, *shape):
            return (g.standard_t(3.0, size=shape) / math.sqrt(3.0)).astype(np.float32)

        for li, (name, cin, cout, k, _s) in enumerate(SEANET_CONVS):
            g = _rng(seed, 11, li)
            gain = 1.0 if li == 0 else 1.45
            if name.endswith("block.3"):
                gain = 0.8
            w = student(g, cout, cin, k) * np.float32(gain / math.sqrt(cin * k))
            if li == 0:
                w *= np.float32(8.0)
            if li in (3, 6, 9):
                w[g.choice(cout, size=3, replace=False)] *= np.float32(50.0)
            if li == 13:
                w *= np.float32(HEAVY_OUT_GAIN)     # bring the SEANet output back to O(1) after the outlier channels
            sd[f"{name}.conv.weight"] = w
        for l in range(N_LAYERS):
            g = _rng(seed, 12, l)
            p = f"encoder_transformer.layers.{l}"
            for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
                sd[f"{p}.self_attn.{nm}.weight"] = student(g, HIDDEN, HIDDEN) * np.float32(1.0 / math.sqrt(HIDDEN))
            sd[f"{p}.self_attn.q_proj.weight"] *= np.float32(2.0)
            w1 = student(g, FFN, HIDDEN) * np.float32(1.0 / math.sqrt(HIDDEN))
            w1[g.choice(FFN, size=4, replace=False)] *= np.float32(50.0)
            sd[f"{p}.mlp.fc1.weight"] = w1
            sd[f"{p}.mlp.fc2.weight"] = student(g, HIDDEN, FFN) * np.float32(1.0 / math.sqrt(FFN))
            for nm in ("self_attn_layer_scale", "mlp_layer_scale"):
                sd[f"{p}.{nm}.scale"] = np.full(HIDDEN, 0.01, np.float32)
        return sd
    raise ValueError(f"unknown weight variant '{kind}'")


# the x50 outlier channels of the "heavy" variant lift the SEANet output rms to ~175 (intermediate activations reach 1e4);
# the last conv is scaled by this constant (measured once, like RVQ_SIGMA0) so that the transformer stream and the RVQ see O(1)
HEAVY_OUT_GAIN = 0.0085
