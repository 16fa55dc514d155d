"""ORACLE (test infrastructure, NOT product code): polyphase FIR resampler, numpy restatement.

The reference resamples with ``librosa.resample(audio, orig_sr, target_sr)`` (REF/*/utils.py:84-87),
whose default backend is libsoxr "HQ" -- a third-party C library that is neither vendored under
/root/reference nor installed here, and whose output no reference test pins.
=> RESAMPLER PARITY UNPINNED. What IS kept from the reference: the call
shape, the no-op when rates are equal, and the output length ceil(n * target/orig) of librosa's
``fix=True``. The filter itself is specified here: a zero-phase Kaiser-windowed sinc applied as a
rational L/M polyphase FIR.

The absent dependency, named: ``librosa.resample(res_type="soxr_hq")`` -> python-soxr (librosa >= 0.10 requires
soxr >= 0.3.2, which bundles libsoxr 0.1.3; the reference pins no version) -> ``soxr_quality_spec(SOXR_HQ)``. Its published
specification (soxr.h / soxr.c): 20-bit precision = 120.4 dB rejection, linear phase, pass band flat up to
1 - 0.05 / TO_3dB(120.4) = 0.9136 of the Nyquist frequency of the lower rate, stop band from 1.0 x Nyquist (no aliasing / imaging
above it), realised as half-band stages plus a polyphase stage with Kaiser-designed taps. THIS filter (response pinned in
tests/test_resample_oracle.py::test_response_against_the_soxr_hq_specification) is SOFTER than that: -0.01 dB at 0.84 x Nyquist,
-0.1 dB at 0.87, -1 dB at 0.90, -1.8 dB at soxr_hq's pass-band edge 0.9136, -3 dB at 0.925, -6 dB at 0.945, only -24 dB at
1.0 x Nyquist, -60 dB at 1.05 and <= -120 dB from 1.08 x Nyquist on. For 16 kHz -> 24 kHz that is a droop of up to 1.8 dB
between 6.7 and 7.3 kHz and partially attenuated images between 8.0 and 8.6 kHz -- a region the Mimi codes are measurably
sensitive to (bench.py --workload c1, resampler_filter_sensitivity). Meeting soxr_hq's mask with one Kaiser stage takes ~2.6 x the
taps (transition 0.087 x Nyquist at 120 dB); the kernel takes the taps as a table, so that is a change of three constants here and in
csrc/mimi_b200.cu: design_taps, not made in this round because it could no longer be validated on a GPU. (soxr's figures are restated from its public
header's description of ``soxr_quality_spec``; the library is not available offline.)

The same taps (computed in float64, rounded to fp32) are what the CUDA kernel uses, so GPU-vs-oracle agreement is at fp32 rounding level; the oracle's indexing is
cross-checked against ``scipy.signal.resample_poly(x, L, M, window=taps)`` and its quality against
``torchaudio.functional.resample`` in tests/test_resample_oracle.py.
"""
import math

import numpy as np

ZEROS = 32          # sinc zero crossings kept on each side (at the narrower of the two rates)
ROLLOFF = 0.945     # cutoff as a fraction of the narrower Nyquist
KAISER_BETA = 14.769656459379492   # ~ -150 dB side lobes


def ratio(sr_in: int, sr_out: int):
    g = math.gcd(int(sr_in), int(sr_out))
    return int(sr_out) // g, int(sr_in) // g          # up L, down M


def design_taps(sr_in: int, sr_out: int):
    """Prototype low-pass on the fine grid (rate L*sr_in), odd length 2c+1, DC gain L.
    Returns (taps float32 [2c+1], L, M, c)."""
    L, M = ratio(sr_in, sr_out)
    fc = ROLLOFF * 0.5 / max(L, M)                   # cycles per fine-grid sample
    c = int(math.ceil(ZEROS / (2.0 * fc)))
    i = np.arange(-c, c + 1, dtype=np.float64)
    h = 2.0 * fc * np.sinc(2.0 * fc * i) * np.kaiser(2 * c + 1, KAISER_BETA)
    h *= L / h.sum()
    return h.astype(np.float32), L, M, c


def out_len(n: int, sr_in: int, sr_out: int) -> int:
    """librosa.resample(fix=True): int(ceil(n * sr_out / sr_in))."""
    return int(math.ceil(n * float(sr_out) / float(sr_in)))


def resample(x: np.ndarray, sr_in: int, sr_out: int) -> np.ndarray:
    """y[m] = sum_n x[n] * h[c + m*M - n*L], m in [0, ceil(n*sr_out/sr_in)); fp32 accumulate in the
    tap order the kernel uses (ascending n)."""
    x = np.asarray(x, np.float32)
    if sr_in == sr_out:                              # REF/*/utils.py:85-86
        return x
    h, L, M, c = design_taps(sr_in, sr_out)
    n = x.shape[0]
    m_out = out_len(n, sr_in, sr_out)
    if n == 0:
        return np.zeros(0, np.float32)
    # zero-stuff to the fine grid, full convolution, pick every M-th sample (fine for test sizes)
    up = np.zeros(n * L, np.float32)
    up[::L] = x
    full = np.convolve(up.astype(np.float64), h.astype(np.float64))       # index i+c <-> fine pos i
    idx = c + np.arange(m_out, dtype=np.int64) * M
    y = np.where(idx < full.shape[0], full[np.minimum(idx, full.shape[0] - 1)], 0.0)
    return y.astype(np.float32)
