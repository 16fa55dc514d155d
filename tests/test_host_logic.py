"""Host-side logic that needs no GPU: feature-extractor mirror, bucketing/sharding, product hygiene."""
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_input, load_golden
from tokenize_audio_b200 import sharding
from tokenize_audio_b200.encoder import EncodecFeatureExtractorLite, MimiEncoderOutput


def test_feature_extractor_mirror_reproduces_reference_padding():
    g = load_golden("mimi_b3_pad_k8")
    lens = g["lengths"].tolist()
    audio = [golden_input(g)[i, 0, :n] for i, n in enumerate(lens)]
    out = EncodecFeatureExtractorLite()(raw_audio=audio, sampling_rate=24000, return_tensors="np", padding=True)
    assert out["input_values"].shape == (3, 1, max(lens)) and out["input_values"].dtype == np.float32
    assert hashlib.sha256(out["input_values"].tobytes()).hexdigest() == str(g["input_values_sha256"])
    assert out["padding_mask"].sum(-1).tolist() == g["padding_mask_sum"].tolist() == lens
    with pytest.raises(ValueError, match="sampling rate"):
        EncodecFeatureExtractorLite()(raw_audio=audio[0], sampling_rate=16000)


def test_single_item_keeps_batch_dim_and_float64_is_downcast():
    out = EncodecFeatureExtractorLite()(raw_audio=np.zeros(100, np.float64), sampling_rate=24000, return_tensors="np")
    assert out["input_values"].shape == (1, 1, 100) and out["input_values"].dtype == np.float32
    assert out["padding_mask"].shape == (1, 100)


def test_encoder_output_is_tuple_compatible():
    o = MimiEncoderOutput("codes")
    assert o.audio_codes == "codes" and o[0] == "codes" and o.padding_cache is None and len(o) == 3


def test_bucket_batches_and_sharding():
    rng = np.random.default_rng(0)
    lens = rng.integers(2 * 24000, 20 * 24000, size=256).tolist()
    batches = sharding.bucket_batches(lens, 64)
    assert sorted(i for b in batches for i in b) == list(range(256))
    assert all(len(b) <= 64 for b in batches)
    file_order = [list(range(i, i + 64)) for i in range(0, 256, 64)]
    assert sharding.padding_waste(lens, batches) < 0.5 * sharding.padding_waste(lens, file_order)
    assert sharding.bucket_batches([], 8) == []
    parts = [sharding.shard_for_rank(10, r, 4) for r in range(4)]
    assert sorted(i for p in parts for i in p) == list(range(10))
    assert sharding.reduce_counters({"a": 1.0, "t_max": 2.0}) == {"a": 1.0, "t_max": 2.0}


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "tokenize_audio_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "import transformers" not in src and "from transformers" not in src, f"{f} imports transformers"


def test_sub_batch_schedule_covers_every_item_once():
    from tokenize_audio_b200.encoder import MimiEncoder
    for B in (2, 3, 8, 9, 16, 24, 64, 100, 257):
        for chunk in (1, 4, 16, 64):
            parts = MimiEncoder._sub_batches(B, chunk)
            flat = [i for p in parts for i in p]
            assert flat == list(range(B))
            assert all(len(p) > 0 for p in parts)
            assert len(parts) <= 3 or chunk < B // 4        # a small head, then at most two big launches
    assert [len(p) for p in MimiEncoder._sub_batches(64, 16)] == [8, 28, 28]


def test_front_groups_partition_the_batch_in_order():
    from tokenize_audio_b200.encoder import MimiEncoder
    for B in (1, 2, 3, 8, 9, 64, 100, 257):
        for first in (1, 4, 8, 1000):
            parts = MimiEncoder._front_groups(B, first)
            assert [i for p in parts for i in p] == list(range(B))
            assert all(len(p) > 0 for p in parts)
            assert all(len(b) <= 2 * len(a) for a, b in zip(parts[:-1], parts[1:]))
    assert [len(p) for p in MimiEncoder._front_groups(64, 8)] == [8, 16, 32, 8]


def test_host_pack_gathers_and_zero_pads():
    """mimi_b200_host_pack (host-only entry point of the C-ABI): ragged clips -> rows of the staging buffer, zeros up to
    zero_to[i], nothing written beyond; any thread count gives the same bytes."""
    import ctypes as C
    import torch
    from tokenize_audio_b200 import _lib
    lib = _lib.load_library()
    rng = np.random.default_rng(3)
    arrs = [rng.standard_normal(int(n)).astype(np.float32) for n in (0, 1, 1919, 1920, 1921, 300000, 77777, 262144, 262145)]
    B, N = len(arrs), max(a.shape[0] for a in arrs)
    zto = [min(N, -(-a.shape[0] // 1920) * 1920) for a in arrs]
    src = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
    lens = (C.c_int64 * B)(*[a.shape[0] for a in arrs])
    z = (C.c_int64 * B)(*zto)
    for threads in (1, 3, 8):
        buf = torch.full((B, 1, N), 7.0)
        assert lib.mimi_b200_host_pack(buf.data_ptr(), buf.stride(0), src, lens, z, B, threads) == 0
        for i, a in enumerate(arrs):
            n = a.shape[0]
            assert np.array_equal(buf[i, 0, :n].numpy(), a)
            assert bool((buf[i, 0, n:zto[i]] == 0).all()) and bool((buf[i, 0, zto[i]:] == 7).all())
    bad = (C.c_int64 * B)(*[N + 1] * B)
    assert lib.mimi_b200_host_pack(buf.data_ptr(), buf.stride(0), src, bad, z, B, 2) != 0


def test_item_ranges_of_equal_work():
    """encode() splits a batch into contiguous item ranges for its side streams: equal total length when the lengths are known
    (every range non-empty, bounds monotone, the last one ends at B), equal item counts otherwise."""
    from tokenize_audio_b200.encoder import MimiB200Model as M
    assert M._range_bounds(64, 2, None) == [0, 32, 64]
    assert M._range_bounds(8, 2, [1, 1, 1, 1, 10, 10, 10, 10]) == [0, 6, 8]
    assert M._range_bounds(8, 2, [100, 1, 1, 1, 1, 1, 1, 1]) == [0, 1, 8]
    assert M._range_bounds(8, 3, [5] * 8) == [0, 3, 5, 8]
    assert M._range_bounds(2, 2, [0, 0]) == [0, 1, 2]
    rng = np.random.default_rng(0)
    for _ in range(50):
        B = int(rng.integers(2, 70))
        n = int(rng.integers(2, min(B, 4) + 1))
        vl = rng.integers(0, 500000, size=B).tolist()
        b = M._range_bounds(B, n, vl)
        assert b[0] == 0 and b[-1] == B and len(b) == n + 1 and all(x < y for x, y in zip(b, b[1:]))
        if n == 2 and sum(vl) > 0:
            # no other split point is closer to half of the work
            half = sum(vl) / 2
            best = min(range(1, B), key=lambda k: abs(sum(max(v, 1) for v in vl[:k]) - sum(max(v, 1) for v in vl) / 2))
            assert abs(sum(vl[:b[1]]) - half) <= abs(sum(vl[:best]) - half) + B


def test_bench_c2_variants_keep_the_items_and_change_only_the_batching():
    """bench.py --order file / --strict: the same 512 utterances, batched in pool order instead of
    by length bucket; the strict variant's config says that every item runs over the padded length."""
    import importlib
    import sys
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        bench = importlib.import_module("bench")
    finally:
        sys.argv = argv
    bench.select_workload("c2")
    keep = bench.ORDER, bench.STRICT, bench.synth.synth_speech
    bench.synth.synth_speech = lambda seed, n, sr=24000: np.zeros(n, np.float32)     # the audio itself is irrelevant here
    try:
        got = {}
        for order in ("bucketed", "file"):
            bench.ORDER = order
            clips, lengths, batches = bench.make_workload(0)
            assert sorted(i for b in batches for i in b) == list(range(bench.POOL))
            assert all(len(b) == bench.BATCH for b in batches)
            assert [[len(c) for c in cl] for cl in clips] == [[lengths[i] for i in b] for b in batches]
            got[order] = sharding.padding_waste(lengths, batches)
        assert batches == [list(range(i, i + bench.BATCH)) for i in range(0, bench.POOL, bench.BATCH)]
        assert got["bucketed"] < 0.2 < 0.4 < got["file"]
        bench.STRICT = True
        assert bench.bench_config()["mode"].startswith("strict") and "file order" in bench.bench_config()["durations_s"]
        bench.ORDER, bench.STRICT = "bucketed", False
        assert bench.bench_config()["mode"].startswith("ragged") and "length-bucketed" in bench.bench_config()["durations_s"]
    finally:
        bench.ORDER, bench.STRICT, bench.synth.synth_speech = keep


def test_resample_audio_defaults_to_the_reference_resampler_when_it_exists(monkeypatch):
    """the drop-in ``resample_audio`` must not silently swap soxr_hq for the GPU filter. backend=None -> librosa
    when importable, else the GPU kernel with ONE RuntimeWarning; explicit names are taken as given; equal rates are a no-op."""
    import importlib.util
    import warnings
    from tokenize_audio_b200 import utils
    x = np.zeros(10, np.float32)
    assert utils.resample_audio(x, 24000, 24000) is x
    assert utils._resampler_backend("b200") == "b200" and utils._resampler_backend("librosa") == "librosa"
    with pytest.raises(ValueError, match="unknown resampler backend"):
        utils._resampler_backend("sox")
    real = importlib.util.find_spec
    monkeypatch.setattr(importlib.util, "find_spec", lambda name, *a, **k: object() if name == "librosa" else real(name, *a, **k))
    assert utils._resampler_backend(None) == "librosa"
    monkeypatch.setattr(importlib.util, "find_spec", lambda name, *a, **k: None if name == "librosa" else real(name, *a, **k))
    monkeypatch.setattr(utils, "_warned_resampler", False)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        assert utils._resampler_backend(None) == "b200" and utils._resampler_backend(None) == "b200"
    assert len([m for m in w if issubclass(m.category, RuntimeWarning)]) == 1
