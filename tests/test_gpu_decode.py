"""Decode direction: MimiB200Model.decode / utils.str_to_audio against waveforms produced by the
real transformers.MimiModel.decode (tests/golden/mimi_decode.npz) and against the oracle. fp32 FFMA path: relative L2 <= 1e-5."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mimi_oracle as O
from tokenize_audio_b200 import synth, utils

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def full_model(state_dict):
    from tokenize_audio_b200.encoder import MimiB200Model
    m = MimiB200Model({**state_dict, **synth.decoder_state_dict(0)}, device="cuda:0")
    yield m
    m.close()


def _rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.mark.parametrize("tag", ["k8", "k32", "k1"])
def test_decode_matches_transformers(full_model, tag):
    g = load_golden("mimi_decode")
    codes = torch.from_numpy(g[f"{tag}_codes"].astype(np.int64)).cuda()
    out = full_model.decode(codes)
    audio = out.audio_values
    assert audio.shape == g[f"{tag}_audio"].shape and audio.dtype == torch.float32 and audio.is_cuda
    assert out[0] is audio and out.decoder_past_key_values is None
    assert _rel(audio.cpu().numpy(), g[f"{tag}_audio"]) <= 1e-5
    tup = full_model.decode(codes, return_dict=False)
    assert isinstance(tup, tuple) and torch.equal(tup[0], audio)
    # a shorter padding mask truncates the waveform (modeling_mimi.py:1668-1670); the decoder is causal
    mask = torch.ones(codes.shape[0], 1000, device="cuda")
    assert torch.equal(full_model.decode(codes, mask).audio_values, audio[..., :1000])
    t = max(1, codes.shape[2] // 2)
    assert _rel(full_model.decode(codes[:, :, :t]).audio_values.cpu().numpy(), g[f"{tag}_audio"][..., : 1920 * t]) <= 1e-5


def test_str_to_audio_and_round_trip(full_model, state_dict):
    g = load_golden("mimi_decode")
    s = g["k8_item0_utf8"].tobytes().decode("utf-8")
    wav = utils.str_to_audio(s, full_model, device="cuda:0")          # REF/emilia-mimi/utils.py:72-81
    assert wav.shape == (1, 1920 * 140) and wav.dtype == np.float32
    assert _rel(wav, g["k8_audio"][0]) <= 1e-5
    # encode -> string -> decode of real (synthetic speech) audio against the oracle on the encoder's own codes
    audio = synth.synth_speech(4100, 24000 * 2 + 300)
    text = utils.audio_to_str(audio, full_model, device="cuda:0")
    codes = np.array(utils.chars_to_codes(text, 8, 2048), np.int64)
    back = utils.str_to_audio(text, full_model, device="cuda:0")
    assert back.shape == (1, 1920 * codes.shape[1])
    sd = {**state_dict, **synth.decoder_state_dict(0)}
    assert _rel(back[None], O.decode(sd, codes[None])) <= 1e-5


def test_decode_argument_checks(full_model, b200_model):
    with pytest.raises(ValueError, match="between 1 and 32 codebooks"):
        full_model.decode(torch.zeros(1, 33, 4, dtype=torch.int64, device="cuda"))
    with pytest.raises(IndexError, match=r"\[0, 2048\)"):
        full_model.decode(torch.full((1, 8, 4), 2048, dtype=torch.int64, device="cuda"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        full_model.decode(torch.zeros(1, 8, 4, dtype=torch.int64))
    with pytest.raises(RuntimeError, match="encode-only"):
        b200_model.decode(torch.zeros(1, 8, 4, dtype=torch.int64, device="cuda"))
    assert full_model.decode(torch.zeros(0, 8, 4, dtype=torch.int64, device="cuda")).audio_values.shape == (0, 1, 7680)
