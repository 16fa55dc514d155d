"""The C-ABI library builds, loads, and exports every symbol include/mimi_b200.h declares. No compute here."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT
from tokenize_audio_b200 import _lib, build


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mimi_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mimi_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = C.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mimi_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding and header disagree"


def test_pure_host_entry_points():
    lib = _lib.load_library()
    assert lib.mimi_b200_abi_version() == 1
    for n, t in [(240777, 126), (150000, 79), (1, 1), (1919, 1), (1920, 1), (1921, 2), (0, 0)]:
        assert lib.mimi_b200_encoded_frames(n) == t
    assert lib.mimi_b200_utf8_bytes_per_frame(8, 0xE000, 2048) == 28
    assert lib.mimi_b200_utf8_bytes_per_frame(32, 0xE000, 2048) == 4 * 3 + 28 * 4
    assert lib.mimi_b200_utf8_bytes_per_frame(32, 0x4E00, 2048) == -1          # surrogates
    assert lib.mimi_b200_resample_out_len(160000, 16000, 24000) == 240000
    assert lib.mimi_b200_resample_out_len(7, 48000, 24000) == 4
    assert C.sizeof(_lib.Weights) == 8 * (14 + 14 + 12 * 8 + 3 + 32 + 32 + 1)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load_library()
    h = C.c_void_p()
    assert lib.mimi_b200_create(C.byref(h), 0) == 2           # MIMI_B200_ERR_CUDA
    assert b"no CPU fallback" in lib.mimi_b200_last_error(None)
    from tokenize_audio_b200.encoder import MimiB200Model
    with pytest.raises(_lib.MimiB200Error):
        MimiB200Model({}, device="cuda")


def test_phase_constants_match_the_header():
    """tokenize_audio_b200/_lib.py mirrors the MIMI_B200_PHASE_* values of include/mimi_b200.h."""
    import os
    import re
    from tokenize_audio_b200 import _lib
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "mimi_b200.h")).read()
    vals = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define MIMI_B200_PHASE_(\w+) (\d+)", hdr)}
    assert vals == {"BEGIN": _lib.PHASE_BEGIN, "FRONT": _lib.PHASE_FRONT, "FINISH": _lib.PHASE_FINISH}
