"""The wrapper's pipelined / storage-format / long-form entry points and the byte kernels behind them (needs a B200).
Everything here is compared with the synchronous reference-shaped call (``encode_audio_batch``), which the parity tests pin
to the oracle and to transformers.MimiModel."""
import threading

import numpy as np
import pytest
import torch

from oracle import chars_oracle as CO
from oracle import mimi_oracle as O
from oracle import resample_oracle as RO
from tokenize_audio_b200 import synth, utils
from tokenize_audio_b200.encoder import MimiEncoder

pytestmark = pytest.mark.gpu


def _clips(seed, lens):
    return [synth.synth_speech(seed + i, n) for i, n in enumerate(lens)]


@pytest.fixture(scope="module")
def batches():
    rng = np.random.default_rng(11)
    return [_clips(2000 + 50 * j, [int(v) for v in rng.uniform(0.5, 6.0, size=b) * 24000]) for j, b in enumerate((12, 1, 9, 16, 3))]


def test_submit_result_and_stream_equal_the_synchronous_call(b200_model, batches):
    enc = MimiEncoder(b200_model, num_quantizers=8)
    want = [enc.encode_audio_batch(b) for b in batches]
    got = list(enc.encode_stream(batches))
    assert len(got) == len(want)
    for w, g in zip(want, got):
        assert len(w) == len(g)
        for a, b in zip(w, g):
            assert a.dtype == b.dtype == np.int64 and np.array_equal(a, b)
    # explicit submit / result, collected out of order, and the depth limit
    p0, p1 = enc.submit(batches[0]), enc.submit(batches[2])
    with pytest.raises(RuntimeError, match="in flight"):
        enc.submit(batches[3])
    r1, r0 = enc.result(p1), enc.result(p0)
    assert all(np.array_equal(a, b) for a, b in zip(r0, want[0])) and all(np.array_equal(a, b) for a, b in zip(r1, want[2]))
    with pytest.raises(RuntimeError, match="already collected"):
        enc.result(p0)
    assert enc.result(enc.submit([])) == []


def test_uint16_and_utf8_formats(b200_model, batches):
    enc = MimiEncoder(b200_model, num_quantizers=8)
    clips = batches[0]
    ref = enc.encode_audio_batch(clips)
    u16 = enc.encode_audio_batch(clips, dtype=np.uint16)      # REF/yodas2-mimi/process_shard.py:519-523
    for a, b in zip(ref, u16):
        assert b.dtype == np.uint16 and np.array_equal(a.astype(np.uint16), b)
    strs = enc.encode_to_strings(clips)
    for a, s in zip(ref, strs):
        assert s.encode("utf-8") == CO.codes_to_utf8(a, 2048)
    tagged = enc.encode_to_strings(clips[:2], audio_tags=("<|audio_start|>", "<|audio_end|>"))
    assert tagged[0] == "<|audio_start|>" + strs[0] + "<|audio_end|>"
    # semantic-only strings = every 8th character of the 8-codebook string (build_yodas2_mm_semantic.py:169-195)
    sem = enc.encode_to_strings(clips, num_codebooks=1)
    assert [s[::8] for s in strs] == sem
    streamed = list(enc.encode_stream([clips, batches[2]], fmt="utf8"))
    assert streamed[0] == strs
    # a 32-codebook wrapper still writes 8-codebook strings by default
    enc32 = MimiEncoder(b200_model)
    assert enc32.encode_to_strings(clips[:3]) == enc.encode_to_strings(clips[:3])


def test_concurrent_callers_share_one_wrapper(b200_model, batches):
    """REF/yodas2-mimi/process_shard.py:690-717 may call the encoder from several threads: the wrapper serialises them."""
    enc = MimiEncoder(b200_model, num_quantizers=8)
    want = [enc.encode_audio_batch(b) for b in batches]
    got, errs = {}, []

    def work(j):
        try:
            for _ in range(3):
                got[j] = enc.encode_audio_batch(batches[j])
        except Exception as e:          # noqa: BLE001
            errs.append(e)
    threads = [threading.Thread(target=work, args=(j,)) for j in range(len(batches))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for j, w in enumerate(want):
        assert all(np.array_equal(a, b) for a, b in zip(w, got[j]))


def test_long_form_audio_as_one_stream_and_as_reference_pieces(b200_model, state_dict):
    """70 s of audio (T25 = 1750: seven attention windows). max_chunk_duration=None is MimiModel.encode on the unsplit signal
    (oracle); max_chunk_duration=30 reproduces the reference's split / encode / concatenate loop bit for bit."""
    audio = synth.synth_speech(3100, 70 * 24000 + 555)
    enc = MimiEncoder(b200_model, num_quantizers=8)
    whole = enc.encode_long_audio(audio)
    ref = O.encode(state_dict, audio[None, None, :], 8)[0]
    assert whole.shape == ref.shape == (8, 876)
    assert (whole == ref).mean() >= 0.999
    pieces = enc.encode_long_audio(audio, max_chunk_duration=30.0)
    cut = 30 * 24000
    manual = np.concatenate([enc.encode_audio_chunk(audio[i:i + cut]) for i in range(0, len(audio), cut)], axis=1)
    assert np.array_equal(pieces, manual)
    assert pieces.shape == (8, 375 + 375 + 126)
    assert np.array_equal(pieces[:, :375], whole[:, :375])            # the first piece has no missing context
    assert not np.array_equal(pieces[:, 375:400], whole[:, 375:400])  # the later ones do (context reset at the cut)


def test_codes_to_chars_rejects_out_of_range_codes():
    bad = np.zeros((8, 4), np.int64)
    bad[3, 2] = 2048
    with pytest.raises(ValueError, match=r"codes must lie in \[0, 2048\)"):
        utils.codes_to_chars(bad, 2048)
    bad[3, 2] = -1
    with pytest.raises(ValueError, match="codes must lie"):
        utils.codes_to_chars(bad, 2048)


def test_codes_to_uint16_any_shape():
    for shape in ((3, 8, 17), (1, 32, 1), (5,), (2, 8, 375)):
        c = torch.randint(0, 2048, shape, device="cuda")
        u = utils.codes_to_uint16(c)
        assert u.dtype == torch.uint16 and u.shape == c.shape
        assert np.array_equal(u.cpu().numpy(), c.cpu().numpy().astype(np.uint16))


@pytest.mark.parametrize("sr_in", [16000, 48000, 8000, 32000, 12000, 44100, 22050])
def test_resampler_kernels_match_the_oracle_on_ragged_batches(sr_in):
    """resample_poly_kernel (16 k, 48 k, 8 k, 32 k, 12 k) and the generic kernel (44.1 k, 22.05 k) against the numpy oracle on
    a ragged batch with odd lengths, including clips shorter than the filter; zero padding behind every item."""
    lens = [sr_in // 2 + 13, 1, 7, 3 * sr_in // 4, 4099]
    clips = [synth.synth_speech(700 + i, n, sr=sr_in) for i, n in enumerate(lens)]
    out, olens = utils.resample_batch(clips, sr_in, 24000)
    assert olens == [RO.out_len(n, sr_in, 24000) for n in lens]
    host = out.cpu().numpy()
    for i, c in enumerate(clips):
        ref = RO.resample(c, sr_in, 24000)
        assert np.abs(host[i, 0, : olens[i]] - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), f"item {i}"
        assert (host[i, 0, olens[i]:] == 0).all()


def test_polyphase_resampler_equals_the_first_draft_kernel():
    """debug knob 16 switches resample() back to the one-thread-per-output kernel: same taps, same sums up to fp32 order."""
    from tokenize_audio_b200 import _lib
    eng = utils._Engine.get("cuda")
    clips = [synth.synth_speech(760 + i, n, sr=16000) for i, n in enumerate((160000, 31999, 5))]
    a, la = utils.resample_batch(clips, 16000, 24000, pad_to=240001)      # odd row stride: the scalar store path
    _lib.check(eng.lib, eng.h, eng.lib.mimi_b200_debug_set(eng.h, 16, 1), "debug_set")
    try:
        b, lb = utils.resample_batch(clips, 16000, 24000, pad_to=240001)
    finally:
        _lib.check(eng.lib, eng.h, eng.lib.mimi_b200_debug_set(eng.h, 16, 0), "debug_set")
    assert la == lb and a.shape == b.shape == (3, 1, 240001)
    assert float((a - b).abs().max()) <= 2e-6


def test_native_rate_front_door(b200_model):
    enc = MimiEncoder(b200_model, num_quantizers=8)
    clips = [synth.synth_speech(780 + i, n, sr=16000) for i, n in enumerate((40000, 16000, 23456))]
    got = enc.encode_native_rate_batch(clips, 16000)
    x24 = [RO.resample(c, 16000, 24000) for c in clips]           # the oracle's 24 kHz samples (same taps as the kernel)
    want = enc.encode_audio_batch(x24)
    assert [g.shape for g in got] == [w.shape for w in want]
    agree = np.mean([np.mean(g == w) for g, w in zip(got, want)])
    assert agree >= 0.99          # inputs agree to fp32 rounding (2e-6), so only near-ties may flip


def test_fp16_range_overflow_falls_back_to_the_tf32_generation(b200_model):
    """The default generation carries activations as fp16 pairs (max 65504). Audio scaled far beyond [-1, 1] drives the
    SEANet activations past that: the kernels clamp and raise the range flag, and the wrapper re-encodes the batch with the
    range-safe TF32 generation -- the result is exactly what a mode-7 wrapper returns, and later batches are unaffected."""
    from tokenize_audio_b200.encoder import MimiB200Model
    loud = [synth.synth_speech(3300 + i, n) * np.float32(3e5) for i, n in enumerate((30000, 20000, 9000))]
    normal = [synth.synth_speech(3310 + i, n) for i, n in enumerate((25000, 12000))]
    b200_model.range_overflow(reset=True)
    b200_model.set_mode(MimiB200Model.RANGE_SAFE_MODE)
    try:
        safe = MimiEncoder(b200_model, num_quantizers=8)
        want_loud, want_normal = safe.encode_audio_batch(loud), safe.encode_audio_batch(normal)
    finally:
        b200_model.set_mode(True)
    enc = MimiEncoder(b200_model, num_quantizers=8)
    got_normal = enc.encode_audio_batch(normal)
    assert enc.range_fallbacks == 0 and not b200_model.range_overflow()
    got = list(enc.encode_stream([loud, normal, normal]))
    assert enc.range_fallbacks >= 1
    assert all(np.array_equal(a, b) for a, b in zip(got[0], want_loud))
    for res in got[1:]:
        for a, b, c in zip(res, want_normal, got_normal):
            assert np.array_equal(a, b) or np.array_equal(a, c)      # redone range-safely (suspect) or plain mode 9
    assert not b200_model.range_overflow()
    after = enc.encode_audio_batch(normal)
    assert all(np.array_equal(a, b) for a, b in zip(after, got_normal))
