"""The oracle (oracle/mimi_oracle.py) against golden vectors produced by the real transformers.MimiModel
(tests/golden/make_golden.py). This is what pins the oracle; the GPU tests then compare CUDA vs oracle."""
import numpy as np
import pytest

from conftest import golden_input, load_golden
from oracle import mimi_oracle as O
from tokenize_audio_b200 import synth


def test_weights_digest_matches_fixture(state_dict):
    g = load_golden("mimi_b1_k32")
    assert synth.state_dict_digest(state_dict) == str(g["weights_digest"]), \
        "synthetic weight generator drifted from the one the golden fixtures were made with"


@pytest.mark.parametrize("name", ["mimi_b1_k32", "mimi_b3_pad_k8", "mimi_long_k8"])
def test_oracle_matches_reference_codes_and_latent(state_dict, name):
    g = load_golden(name)
    x = golden_input(g)
    K = int(g["num_quantizers"])
    taps = {}
    codes = O.encode(state_dict, x, K, taps=taps)
    ref = g["codes"].astype(np.int64)
    assert codes.shape == ref.shape and codes.dtype == np.int64
    agree = float((codes == ref).mean())
    assert agree >= 0.999, f"oracle/reference code agreement {agree:.5f} < 99.9 %"
    lat = np.stack(taps["latent"])
    rel = np.linalg.norm(lat - g["latent"]) / np.linalg.norm(g["latent"])
    assert rel < 1e-5, f"latent relative L2 error {rel:.2e}"       # fp32 summation-order noise only
    np.testing.assert_allclose(np.stack(taps["seanet.out"])[0][:, :8], g["seanet_out_first"], rtol=0, atol=2e-4)


@pytest.mark.parametrize("name,variant", [("mimi_c3_k8", None), ("mimi_c4_k32", None), ("mimi_ties_k32", "ties"),
                                          ("mimi_heavy_k32", "heavy")])
def test_oracle_matches_reference_at_full_sizes_and_on_adversarial_weights(state_dict, name, variant):
    """BASELINE configs 3 and 4 at full length (30 s K=8, 15 s K=32), a codebook full of exact ties, and a heavy-tailed
    second weight draw: the oracle against the real transformers.MimiModel outputs."""
    g = load_golden(name)
    sd = synth.variant_state_dict(variant) if variant else state_dict
    assert synth.state_dict_digest(sd) == str(g["weights_digest"])
    x = golden_input(g)
    K = int(g["num_quantizers"])
    taps = {}
    codes = O.encode(sd, x, K, taps=taps)
    ref = g["codes"].astype(np.int64)
    assert (codes == ref).mean() >= 0.999
    for b, k, t in np.argwhere(codes != ref):
        first_bad = int(np.argmax(codes[b, :, t] != ref[b, :, t]))
        assert k > first_bad or g["margins"][b, k, t] < 1e-3
    lat = np.stack(taps["latent"])
    assert np.linalg.norm(lat - g["latent"]) / np.linalg.norm(g["latent"]) < 1e-5
    if variant == "ties":
        assert ref.max() < 1024                       # every winner has a twin at index + 1024: the lower one is returned
        assert (ref[:, [4, 8] + list(range(21, 32))] == 5).mean() > 0.9     # all-zero rows 5, 700, 1029, 1724: index 5
        assert codes.max() < 1024


def test_encoded_length_known_answers():
    g = load_golden("encoded_length")
    for n, t in zip(g["lengths"].tolist(), g["frames"].tolist()):
        assert O.encoded_length(n) == t == -(-n // 1920)


def test_padding_mask_is_ignored_and_k_prefix(state_dict):
    x = synth.synth_speech(5, 5000)[None, None, :]
    full = O.encode(state_dict, x, 32)
    k8 = O.encode(state_dict, x, 8, padding_mask=np.zeros((1, 5000), np.int32))
    assert np.array_equal(full[:, :8], k8)          # first 8 rows of a 32-codebook encode == 8-codebook encode


def test_causality_prefix_property(state_dict):
    x = synth.synth_speech(6, 3 * 1920 + 500)[None, None, :]
    full = O.encode(state_dict, x, 4)
    pre = O.encode(state_dict, x[:, :, : 2 * 1920], 4)
    assert np.array_equal(full[:, :, :2], pre)      # a prefix of whole frames gives the same codes


def test_error_behaviour(state_dict):
    x = np.zeros((1, 1, 1920), np.float32)
    with pytest.raises(ValueError, match="lower than the total number of quantizers 32"):
        O.encode(state_dict, x, 33)
    with pytest.raises(ValueError, match="higher than the number of semantic quantizers 1"):
        O.encode(state_dict, x, 0)
    with pytest.raises(ValueError, match="channels must be 1 or 2"):
        O.encode(state_dict, np.zeros((1, 3, 1920), np.float32), 8)


@pytest.mark.parametrize("tag", ["k8", "k32", "k1"])
def test_oracle_decode_matches_transformers(state_dict, tag):
    """Decode direction: the oracle's MimiModel.decode restatement against waveforms from the real transformers model."""
    g = load_golden("mimi_decode")
    sd = {**state_dict, **synth.decoder_state_dict(0)}
    assert synth.state_dict_digest(sd) == str(g["weights_digest"])
    codes = g[f"{tag}_codes"].astype(np.int64)
    if tag == "k8":
        codes = codes[:1, :, :40]            # the oracle is exact-causal: a prefix of frames gives a prefix of samples
    audio = O.decode(sd, codes)
    ref = g[f"{tag}_audio"][: codes.shape[0], :, : 1920 * codes.shape[2]]
    assert audio.shape == ref.shape and audio.dtype == np.float32
    assert np.linalg.norm(audio - ref) / np.linalg.norm(ref) < 1e-5
