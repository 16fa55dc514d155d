"""The default CUDA path against outputs of the REAL transformers.MimiModel at BASELINE.json's full sizes and on
adversarial weights (fixtures made by tests/golden/make_golden.py; needs a B200):

* C3: two ragged long-form items, 30 s and 22 s (T25 = 750 = three 250-frame attention windows), 8 codebooks;
* C4: 2 x 15 s (360 000 samples, partial last frame), all 32 codebooks -- the residual chain amplifies error;
* a codebook full of EXACT ties (duplicated rows, all-zero dead rows that win): lowest index among equal minima
  (modeling_mimi.py:1200-1201);
* a second, heavy-tailed weight draw (Student-t weights, LayerScale 0.01, x50 outlier channels, activations up to 1e4);
* num_quantizers = 1 (the semantic-only consumer, REF/yodas2-mimi/build_yodas2_mm_semantic.py:169-195);
* utils.audio_to_str (REF/emilia-mimi/utils.py:58-69).

Tolerances: codes identical on >= 99.9 % of slots, every mismatch a near-tie of the reference's own top-2 margin
(< 1e-3, stored in the fixture) or a cascade of one; pre-quantisation latent relative L2 <= 2e-5.
"""
import numpy as np
import pytest
import torch

from conftest import golden_input, load_golden
from tokenize_audio_b200 import synth

pytestmark = pytest.mark.gpu

CODE_AGREEMENT = 0.999
LATENT_REL_TOL = 2e-5
NEAR_TIE = 1e-3


def _rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _check(model, g, ragged):
    x = golden_input(g)
    K = int(g["num_quantizers"])
    lens = g["lengths"].tolist()
    xd = torch.from_numpy(x).to(model.device)
    out, lat = model.encode(xd, num_quantizers=K, return_latent=True, valid_lengths=lens if ragged else None)
    codes, lat = out.audio_codes.cpu().numpy(), lat.cpu().numpy()
    ref = g["codes"].astype(np.int64)
    assert codes.shape == ref.shape and codes.dtype == np.int64
    flips = []
    for i, n in enumerate(lens):
        t = -(-n // 1920) if ragged else ref.shape[2]          # ragged mode only promises the kept frames
        assert _rel(lat[i, :, :t], g["latent"][i, :, :t]) <= LATENT_REL_TOL, f"item {i}"
        bad = codes[i, :, :t] != ref[i, :, :t]
        assert 1.0 - bad.mean() >= CODE_AGREEMENT, f"item {i}: agreement {1.0 - bad.mean():.5f}"
        for k, f in np.argwhere(bad):
            first_bad = int(np.argmax(bad[:, f]))
            assert k > first_bad or g["margins"][i, k, f] < NEAR_TIE, \
                f"unexplained flip item {i} cb {k} frame {f} margin {g['margins'][i, k, f]:.2e}"
            flips.append((i, int(k), int(f)))
    return codes, flips


@pytest.fixture(scope="module")
def variant_models():
    from tokenize_audio_b200.encoder import MimiB200Model
    cache = {}

    def get(kind):
        if kind not in cache:
            cache[kind] = MimiB200Model(synth.variant_state_dict(kind), device="cuda:0")
        return cache[kind]
    yield get
    for m in cache.values():
        m.close()


@pytest.mark.parametrize("ragged", [False, True])
@pytest.mark.parametrize("name", ["mimi_c3_k8", "mimi_c4_k32"])
def test_full_size_configs_against_transformers(b200_model, name, ragged):
    _check(b200_model, load_golden(name), ragged)


def test_exact_ties_take_the_lowest_index(variant_models):
    g = load_golden("mimi_ties_k32")
    codes, _ = _check(variant_models("ties"), g, ragged=False)
    assert codes.max() < 1024, "a duplicated row at index + 1024 won an exact tie"
    zero_stages = [4, 8] + list(range(21, 32))                  # acoustic stages 3, 7, 20..30: all-zero rows 5, 700, 1029, 1724
    ref = g["codes"].astype(np.int64)
    assert np.array_equal(codes[:, zero_stages] == 5, ref[:, zero_stages] == 5)
    assert (codes[:, zero_stages] == 5).mean() > 0.9
    assert not np.isin(codes[:, zero_stages], [700, 1029, 1724]).any()


@pytest.mark.parametrize("ragged", [False, True])
def test_heavy_tailed_weights(variant_models, ragged):
    _check(variant_models("heavy"), load_golden("mimi_heavy_k32"), ragged)


@pytest.mark.parametrize("name", ["mimi_b3_pad_k8", "mimi_c3_k8"])
def test_semantic_only_num_quantizers_1(b200_model, name):
    """K = 1: the semantic codebook alone; the acoustic projection and chain are skipped, the row equals row 0 of a
    larger encode (MimiSplitResidualVectorQuantizer.encode, modeling_mimi.py:1311-1338)."""
    g = load_golden(name)
    xd = torch.from_numpy(golden_input(g)).to(b200_model.device)
    k1 = b200_model.encode(xd, num_quantizers=1).audio_codes
    assert k1.shape == (xd.shape[0], 1, g["codes"].shape[2]) and k1.dtype == torch.int64
    ref = g["codes"].astype(np.int64)[:, :1]
    assert (k1.cpu().numpy() == ref).mean() >= CODE_AGREEMENT
    k8 = b200_model.encode(xd, num_quantizers=8).audio_codes
    assert torch.equal(k1, k8[:, :1])


def test_audio_to_str_against_reference_pipeline(b200_model):
    from tokenize_audio_b200 import utils
    g = load_golden("audio_to_str")
    audio = g["pcm"].astype(np.float32) / np.float32(32768.0)
    want = g["utf8"].tobytes().decode("utf-8")
    got = utils.audio_to_str(audio, b200_model, device="cuda:0")
    assert isinstance(got, str) and len(got) == len(want) == 8 * (-(-len(audio) // 1920))
    same = sum(a == b for a, b in zip(got, want))
    assert same / len(want) >= CODE_AGREEMENT
    # the [1, N] form the reference also accepts (it unsqueezes to [1, 1, N] itself)
    assert utils.audio_to_str(audio[None, :], b200_model, device="cuda:0") == got
