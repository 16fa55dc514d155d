"""The kernel generations of the CUDA path against the oracle and against each other, and scheduling invariants (needs a B200).

mode 0 = all-fp32 FFMA (the exact-fp32 baseline), 7 = fused front end + CTA-pair tcgen05 GEMM with TF32 hi / bf16 lo operands +
tcgen05 attention + tensor-core RVQ (fp32 range: the fallback of the default), 9 = default: the same with every GEMM operand
as an fp16 hi/lo pair, all products on kind::f16.
Tolerances as in test_gpu_parity.py: codes >= 99.9 % identical to the oracle, latent relative L2 <= 2e-5.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import chars_oracle as CO
from oracle import mimi_oracle as O
from tokenize_audio_b200 import _lib, synth

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


@pytest.fixture(scope="module")
def case(state_dict):
    # two ragged items, T25 = 302 > 250-frame window on the long one; 32 codebooks
    lens = [289234, 61111]
    x = np.zeros((2, 1, lens[0]), np.float32)
    for i, n in enumerate(lens):
        x[i, 0, :n] = synth.synth_speech(800 + i, n)
    taps = {}
    ref = O.encode(state_dict, x, 32, taps=taps)
    return x, lens, ref, np.stack(taps["latent"])


@pytest.mark.parametrize("mode", [0, 7, 9])
def test_every_mode_matches_the_oracle(b200_model, case, mode):
    x, lens, ref, lat_ref = case
    b200_model.set_mode(mode)
    try:
        out, lat = b200_model.encode(torch.from_numpy(x).cuda(), num_quantizers=32, return_latent=True)
        codes = out.audio_codes.cpu().numpy()
        rag = b200_model.encode(torch.from_numpy(x).cuda(), num_quantizers=32, valid_lengths=lens).audio_codes.cpu().numpy()
    finally:
        b200_model.set_mode(True)
    assert _rel(lat.cpu().numpy(), lat_ref) <= 2e-5
    assert (codes == ref).mean() >= 0.999, f"mode {mode}: {(codes == ref).mean():.5f}"
    for i, n in enumerate(lens):
        t = -(-n // 1920)
        assert np.array_equal(rag[i, :, :t], codes[i, :, :t])          # ragged == strict on kept frames, every mode


def test_front_end_and_attention_schedule_variants_agree(b200_model, case):
    """The fp16 front end (front_f16.cuh, default) against round 1's front end with TF32 internals (debug knob 17), and the
    attention kernel's compact reversed unit list against its grid walk (knob 18; ragged calls only): same codes, latents
    within the tolerance of either against the oracle. The 13-sample and 127-sample items put item ends inside the first
    front-end tile, right before a tile boundary (126) and right after it."""
    x, lens, ref, lat_ref = case
    xd = torch.from_numpy(x).cuda()
    out = {}
    for name, knob in (("default", None), ("front_tf32", 17), ("att_grid", 18)):
        if knob is not None:
            b200_model.debug_set(knob, 1)
        try:
            o, lat = b200_model.encode(xd, num_quantizers=32, return_latent=True)
            rag = b200_model.encode(xd, num_quantizers=32, valid_lengths=lens).audio_codes.cpu().numpy()
            out[name] = (o.audio_codes.cpu().numpy(), lat.cpu().numpy(), rag)
        finally:
            if knob is not None:
                b200_model.debug_set(knob, 0)
    for name, (codes, lat, rag) in out.items():
        assert _rel(lat, lat_ref) <= 2e-5, name
        assert (codes == ref).mean() >= 0.999, name
        for i, n in enumerate(lens):
            t = -(-n // 1920)
            assert np.array_equal(rag[i, :, :t], codes[i, :, :t]), name
    assert np.array_equal(out["default"][2], out["att_grid"][2])            # the schedule changes no arithmetic
    assert (out["default"][0] == out["front_tf32"][0]).mean() >= 0.999
    # tile-boundary lengths of the front end's item-major tile walk, ragged against strict
    short = [13, 126, 127, 1920 * 2 + 5, 252, 1]
    n = max(short)
    xs = np.zeros((len(short), 1, n), np.float32)
    for i, m in enumerate(short):
        xs[i, 0, :m] = synth.synth_speech(850 + i, m)
    a = b200_model.encode(torch.from_numpy(xs).cuda(), num_quantizers=8).audio_codes.cpu().numpy()
    r = b200_model.encode(torch.from_numpy(xs).cuda(), num_quantizers=8, valid_lengths=short).audio_codes.cpu().numpy()
    for i, m in enumerate(short):
        t = -(-m // 1920)
        assert np.array_equal(r[i, :, :t], a[i, :, :t])


def test_tap_group_gemm_matches_per_tap_gemm(b200_model, case):
    """Convs run as tap groups (tc_gemm7.cuh: one activation tile per group of taps that share input rows, shifted
    descriptors) against the per-k-block schedule of tc_gemm5.cuh (debug knob 20): both within tolerance of the oracle, the
    same codes up to near-ties, ragged == strict on kept frames."""
    x, lens, ref, lat_ref = case
    xd = torch.from_numpy(x).cuda()
    out, lat = b200_model.encode(xd, num_quantizers=32, return_latent=True)
    a = out.audio_codes.cpu().numpy()
    ar = b200_model.encode(xd, num_quantizers=32, valid_lengths=lens).audio_codes.cpu().numpy()
    b200_model.debug_set(20, 1)
    try:
        out_b, lat_b = b200_model.encode(xd, num_quantizers=32, return_latent=True)
        b = out_b.audio_codes.cpu().numpy()
    finally:
        b200_model.debug_set(20, 0)
    assert _rel(lat.cpu().numpy(), lat_ref) <= 2e-5 and _rel(lat_b.cpu().numpy(), lat_ref) <= 2e-5
    assert _rel(lat.cpu().numpy(), lat_b.cpu().numpy()) <= 1e-5
    assert (a == ref).mean() >= 0.999 and (a == b).mean() >= 0.999
    for i, n in enumerate(lens):
        t = -(-n // 1920)
        assert np.array_equal(ar[i, :, :t], a[i, :, :t])


def test_fp16_rvq_matches_tf32_rvq(b200_model, case):
    """The fp16-pair RVQ kernel (rvq_f16.cuh, default of generation 9) against the TF32 one (debug knob 19) on the same latents:
    32 codebooks, strict and ragged; the codes may differ only through near-ties (<= 0.1 % of slots) and never on the first
    codebook of this case."""
    x, lens, ref, _ = case
    xd = torch.from_numpy(x).cuda()
    a = b200_model.encode(xd, num_quantizers=32).audio_codes.cpu().numpy()
    ar = b200_model.encode(xd, num_quantizers=32, valid_lengths=lens).audio_codes.cpu().numpy()
    b200_model.debug_set(19, 1)
    try:
        b = b200_model.encode(xd, num_quantizers=32).audio_codes.cpu().numpy()
    finally:
        b200_model.debug_set(19, 0)
    assert (a == b).mean() >= 0.999 and np.array_equal(a[:, 0], b[:, 0])
    assert (a == ref).mean() >= 0.999
    for i, n in enumerate(lens):
        t = -(-n // 1920)
        assert np.array_equal(ar[i, :, :t], a[i, :, :t])


def test_tensor_core_rvq_matches_simt_rvq(b200_model):
    """mode 0 (fused SIMT RVQ, exact fp32 FFMA distances, on the all-fp32 pipeline) vs the default generation (tensor-core
    RVQ): the codes may differ only through near-ties (<= 0.1 % of slots)."""
    x = np.stack([synth.synth_speech(900 + i, 20 * 1920) for i in range(6)])[:, None, :]
    xd = torch.from_numpy(x).cuda()
    b200_model.set_mode(0)
    try:
        a = b200_model.encode(xd, num_quantizers=32).audio_codes
    finally:
        b200_model.set_mode(True)
    b = b200_model.encode(xd, num_quantizers=32).audio_codes
    assert a.shape == b.shape == (6, 32, 20)
    assert float((a == b).float().mean()) >= 0.999
    assert torch.equal(a[:, 0], b[:, 0])


def test_wrapper_staging_schedules_are_invisible(b200_model):
    """encode_audio_batch (ragged mode) returns the same codes whether the batch is staged group by group under the
    running front end (mimi_b200_encode_phase, any group sizes), as independent sub-batches, or in one plain call."""
    from tokenize_audio_b200.encoder import MimiEncoder
    rng = np.random.default_rng(7)
    clips = [synth.synth_speech(1000 + i, int(n)) for i, n in enumerate(rng.integers(3000, 60000, size=21))]
    n = max(len(c) for c in clips)
    x = np.zeros((21, 1, n), np.float32)
    for i, c in enumerate(clips):
        x[i, 0, : len(c)] = c
    plain = b200_model.encode(torch.from_numpy(x).cuda(), num_quantizers=8, valid_lengths=[len(c) for c in clips]).audio_codes.cpu().numpy()
    results = []
    for first in (1, 4, 64):
        results.append(MimiEncoder(b200_model, num_quantizers=8, first_items=first).encode_audio_batch(clips))
    for chunk in (64, 4):
        w = MimiEncoder(b200_model, num_quantizers=8, chunk_items=chunk)
        w.phased = False
        results.append(w.encode_audio_batch(clips))
    for res in results:
        assert len(res) == 21
        for i, (a, c) in enumerate(zip(res, clips)):
            t = -(-len(c) // 1920)
            assert a.shape == (8, t) and np.array_equal(a, plain[i, :, :t])


def test_config5_codes_to_unicode_after_encode(b200_model):
    """BASELINE config 5: 30 s segments -> 8 codebooks -> unicode string (375 frames = 3000 chars = 10 500 bytes)."""
    from tokenize_audio_b200 import utils
    x = torch.from_numpy(np.stack([synth.synth_speech(1100 + i, 720000) for i in range(2)])[:, None, :]).cuda()
    codes = b200_model.encode(x, num_quantizers=8).audio_codes
    got = utils.codes_to_utf8_batch(codes, [375, 375])
    host = codes.cpu().numpy()
    for i in range(2):
        assert len(got[i]) == 10500 and got[i] == CO.codes_to_utf8(host[i], 2048)
        assert len(got[i].decode("utf-8")) == 3000


def test_sw128_descriptor_row_shift_property():
    """The hardware property front_fused.cuh relies on: a SWIZZLE_128B K-major operand descriptor may
    start at any whole 128-byte row of a staged tile when its base-offset field is 0."""
    lib = _lib.load_library()
    h = C.c_void_p()
    _lib.check(lib, None, lib.mimi_b200_create(C.byref(h), 0), "create")
    try:
        g = torch.Generator().manual_seed(1)
        K = 96
        a = torch.randn(136, K, generator=g)
        w = torch.randn(64, K, generator=g)
        ad = a.cuda()
        wn = np.ascontiguousarray(w.numpy())
        for shift in (0, 1, 2, 3, 7, 8):
            out = torch.zeros(128, 64, device="cuda")
            rc = lib.mimi_b200_debug_shift_probe(h, ad.data_ptr(), wn.ctypes.data, K, shift, 0, out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream)
            _lib.check(lib, h, rc, "shift_probe")
            ref = a[shift:shift + 128].double() @ w.double().T
            err = float((out.cpu().double() - ref).norm() / ref.norm())
            assert err < 2e-3, f"shift {shift}: {err:.2e}"          # single-pass TF32
    finally:
        lib.mimi_b200_destroy(h)


def test_front_door_and_string_outputs(b200_model):
    """native-rate clips resampled on the GPU then encoded == resample_audio + encode;
    encode_to_strings == encode + codes_to_chars per item."""
    from tokenize_audio_b200 import utils
    from tokenize_audio_b200.encoder import MimiEncoder
    enc = MimiEncoder(b200_model, num_quantizers=8)
    clips16 = [synth.synth_speech(1200 + i, n, sr=16000) for i, n in enumerate([16000, 23456, 9000])]
    got = enc.encode_native_rate_batch(clips16, 16000)
    clips24 = [utils.resample_audio(c, 16000, 24000, backend="b200") for c in clips16]
    want = enc.encode_audio_batch(clips24)
    assert [g.shape for g in got] == [w.shape for w in want]
    same = sum(int((g == w).sum()) for g, w in zip(got, want)) / sum(w.size for w in want)
    assert same >= 0.999          # same kernel, same filter: only the batch padding differs
    strs = enc.encode_to_strings(clips24)
    for s_, w in zip(strs, want):
        assert s_ == utils.codes_to_chars(w, 2048)
        assert len(s_) == w.size
    assert enc.encode_native_rate_batch([], 16000) == [] and enc.encode_to_strings([]) == []


def test_item_ranges_on_two_streams_are_bit_identical(b200_model):
    """encode() splits a batch of >= 8 items into two contiguous item ranges on side streams (the other range's kernels
    fill the SMs a persistent grid's last tiles leave idle). Items are independent and a tile's arithmetic does not depend
    on which items share its CTA pair, so codes and latents are bit-identical to the single-stream call."""
    rng = np.random.default_rng(11)
    lens = [int(v) for v in rng.integers(2000, 90000, size=13)]
    x = np.zeros((13, 1, max(lens)), np.float32)
    for i, n in enumerate(lens):
        x[i, 0, :n] = synth.synth_speech(1700 + i, n)
    xd = torch.from_numpy(x).cuda()
    res = {}
    keep = b200_model.streams
    try:
        for n_streams in (1, 2):
            b200_model.streams = n_streams
            for ragged in (False, True):
                out, lat = b200_model.encode(xd, num_quantizers=32, valid_lengths=lens if ragged else None, return_latent=True)
                torch.cuda.synchronize()
                res[(n_streams, ragged)] = (out.audio_codes.cpu().numpy(), lat.cpu().numpy())
    finally:
        b200_model.streams = keep
    for ragged in (False, True):
        for i, n in enumerate(lens):
            t = -(-n // 1920) if ragged else res[(1, ragged)][0].shape[2]
            assert np.array_equal(res[(1, ragged)][0][i, :, :t], res[(2, ragged)][0][i, :, :t])
            assert np.array_equal(res[(1, ragged)][1][i, :, :t], res[(2, ragged)][1][i, :, :t])


def test_nothing_is_read_before_it_is_written(b200_model):
    """The whole workspace is filled with NaN before each call: a plain encode, the phased wrapper (several front-end
    groups) and the independent sub-batches must still return the clean result -- i.e. no kernel reads an activation,
    halo row, length or tile-list entry that this call has not written. (Caught a wrong bf16 `lo` offset of the per-group
    front-end launches that stale data from earlier calls had been masking.)"""
    from tokenize_audio_b200.encoder import MimiEncoder
    rng = np.random.default_rng(7)
    clips = [synth.synth_speech(1000 + i, int(n)) for i, n in enumerate(rng.integers(3000, 60000, size=21))]
    lens = [len(c) for c in clips]
    x = np.zeros((21, 1, max(lens)), np.float32)
    for i, c in enumerate(clips):
        x[i, 0, : len(c)] = c
    xd = torch.from_numpy(x).cuda()
    clean = b200_model.encode(xd, num_quantizers=8, valid_lengths=lens).audio_codes.cpu().numpy()

    def poison():
        for w in b200_model._workspaces.values():
            if w is not None:
                w.view(torch.float32)[: w.numel() // 4].fill_(float("nan"))
        torch.cuda.synchronize()

    poison()
    assert np.array_equal(b200_model.encode(xd, num_quantizers=8, valid_lengths=lens).audio_codes.cpu().numpy(), clean)
    poison()
    strict = b200_model.encode(xd, num_quantizers=8).audio_codes.cpu().numpy()
    for i, n in enumerate(lens):
        assert np.array_equal(strict[i, :, : -(-n // 1920)], clean[i, :, : -(-n // 1920)])
    for kw, phased in ((dict(first_items=1), True), (dict(first_items=4), True), (dict(chunk_items=4), False)):
        w = MimiEncoder(b200_model, num_quantizers=8, **kw)
        w.phased = phased
        poison()
        res = w.encode_audio_batch(clips)
        for i, a in enumerate(res):
            assert np.array_equal(a, clean[i, :, : a.shape[1]]), f"{kw}: item {i}"
