"""ORACLE (test infrastructure, NOT product code): polyphase FIR resampler, numpy restatement.

The reference resamples with ``librosa.resample(audio, orig_sr, target_sr)`` (REF/*/utils.py:84-87),
whose default backend is libsoxr "HQ" -- a third-party C library that is neither vendored under
/root/reference nor installed here, and whose output no reference test pins.
=> RESAMPLER PARITY UNPINNED. What IS kept from the reference: the call
shape, the no-op when rates are equal, and the output length ceil(n * target/orig) of librosa's
``fix=True``. The filter itself is specified here: a zero-phase Kaiser-windowed sinc applied as a
rational L/M polyphase FIR. The same taps (computed in float64, rounded to fp32) are what the CUDA
kernel uses, so GPU-vs-oracle agreement is at fp32 rounding level; the oracle's indexing is
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
