"""codes -> unicode oracle against vectors made by the reference's own converter.py."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import chars_oracle as CO


@pytest.mark.parametrize("tag,K,T", [("k8_t375", 8, 375), ("k8_t3", 8, 3), ("k32_t17", 32, 17), ("k1_t5", 1, 5), ("k8_t0", 8, 0)])
def test_utf8_matches_reference_converter(tag, K, T):
    g = load_golden("codes_to_chars")
    codes = g[f"{tag}_codes"].astype(np.int64)
    assert codes.shape == (K, T)
    got = CO.codes_to_utf8(codes, 2048, 0xE000)
    assert got == g[f"{tag}_utf8"].tobytes()
    s = got.decode("utf-8")
    assert len(s) == K * T
    if T:
        assert np.array_equal(CO.chars_to_codes(s, K, 2048, 0xE000), codes)
        assert CO.codes_to_chars(codes, 2048) == s


def test_28_bytes_per_frame_and_layout():
    codes = np.arange(24, dtype=np.int64).reshape(8, 3)
    b = CO.codes_to_utf8(codes, 2048)
    assert len(b) == 28 * 3                       # cb 0-3 -> 3 bytes, cb 4-7 -> 4 bytes
    s = b.decode("utf-8")
    assert [ord(c) for c in s[:8]] == [0xE000 + k * 2048 + codes[k, 0] for k in range(8)]


def test_surrogate_guard_and_shape_errors():
    with pytest.raises(ValueError):
        CO.validate_unicode_offset(0x4E00, 32, 2048)      # runs into U+D800
    assert CO.validate_unicode_offset(0xE000, 32, 2048) == 0xE000
    with pytest.raises(ValueError, match="2D array"):
        CO.codes_to_codepoints(np.zeros((2, 2, 2), np.int64), 2048)
