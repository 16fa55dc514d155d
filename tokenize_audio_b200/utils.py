"""Drop-in counterparts of the reference's ``utils.py`` (REF/emilia-mimi/utils.py, six identical copies)
with the array work done by libmimi_b200.so kernels: ``codes_to_chars``, ``chars_to_codes``,
``audio_to_str``, ``str_to_audio``, ``resample_audio``. Same names, argument meaning and error behaviour.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

UNICODE_OFFSET: int = 0xE000      # REF/emilia-mimi/utils.py:13-15
NUM_CODEBOOKS: int = 8
CODEBOOK_SIZE: int = 2048


class _Engine:
    """Weight-less engine handle for the resampler / UTF-8 / pack kernels (one per device, shared by every caller).

    The C handle is not re-entrant (include/mimi_b200.h): ``lock`` serialises the host side of every call on it, and the
    library itself orders the device-side reuse of its small length tables across streams with an event. The pinned /
    device staging buffers of :func:`resample_batch` live here too, so that no call pins fresh host memory."""

    _cache = {}
    _cache_lock = threading.Lock()

    def __init__(self, device: torch.device):
        self.lib = _lib.load_library()
        self.device = device
        self.lock = threading.RLock()
        h = C.c_void_p()
        with torch.cuda.device(device):
            _lib.check(self.lib, None, self.lib.mimi_b200_create(C.byref(h), device.index), "mimi_b200_create")
        self.h = h
        self._pinned_in: Optional[torch.Tensor] = None
        self._dev_in: Optional[torch.Tensor] = None
        self._staged = None          # event: the last H2D copy out of _pinned_in has completed

    def staging(self, n: int) -> "tuple[torch.Tensor, torch.Tensor]":
        """Persistent pinned + device fp32 staging buffers of at least ``n`` samples (call with ``lock`` held)."""
        if self._pinned_in is None or self._pinned_in.numel() < n:
            self._pinned_in = self._dev_in = None
            cap = max(int(n * 1.25), 1 << 16)
            self._pinned_in = torch.empty(cap, dtype=torch.float32).pin_memory()
            self._dev_in = torch.empty(cap, dtype=torch.float32, device=self.device)
            self._staged = None
        if self._staged is not None:
            self._staged.synchronize()          # the previous call's H2D copy has left the pinned buffer
        return self._pinned_in, self._dev_in

    @classmethod
    def get(cls, device=None) -> "_Engine":
        if not torch.cuda.is_available():
            raise _lib.MimiB200Error("a CUDA device (B200) is required; there is no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
        with cls._cache_lock:
            if dev not in cls._cache:
                cls._cache[dev] = cls(dev)
            return cls._cache[dev]


def validate_unicode_offset(unicode_offset: int, num_codebooks: int, codebook_size: int) -> int:
    """REF/pretraining-data/converter.py:68-81."""
    lower, upper = unicode_offset, unicode_offset + num_codebooks * codebook_size
    if lower < 0xDFFF and upper > 0xD800:
        raise ValueError(
            f"You are using unicode offset {hex(unicode_offset)}, however your base vocabulary size (num_codebooks x codebook_size) "
            f"is {num_codebooks * codebook_size} which will intersect with the non-printable surrogate range 0xD800-0xDFFF if starting from this offset.\n"
            f"To avoid this issue, use a unicode offset starting after the surrogate range, such as {hex(0xE000)}.")
    return unicode_offset


def codes_to_utf8_device(codes: torch.Tensor, frames: Optional[Sequence[int]] = None, codebook_size: int = CODEBOOK_SIZE,
                         unicode_offset: int = UNICODE_OFFSET) -> "tuple[torch.Tensor, List[int]]":
    """codes ``[B,K,T]`` int64 on the GPU -> (``[B, stride]`` uint8 device tensor of UTF-8 rows, byte length of each row);
    asynchronous on the current stream. The codes must lie in ``[0, codebook_size)``."""
    if codes.dim() != 3:
        raise ValueError("codes must be [batch, num_codebooks, seq_length]")
    B, K, T = codes.shape
    validate_unicode_offset(unicode_offset, K, codebook_size)
    eng = _Engine.get(codes.device)
    bpf = eng.lib.mimi_b200_utf8_bytes_per_frame(K, unicode_offset, codebook_size)
    if bpf < 0:
        raise ValueError("unicode offset / codebook size not representable as fixed-width UTF-8 per codebook")
    codes = codes.to(torch.int64).contiguous()
    stride = max(int(T * bpf), 1)
    out = torch.empty((B, stride), dtype=torch.uint8, device=codes.device)
    if B == 0:
        return out, []
    fr = (C.c_int64 * B)(*[int(f) for f in frames]) if frames is not None else None
    lens = (C.c_int64 * B)()
    with eng.lock, torch.cuda.device(codes.device):
        rc = eng.lib.mimi_b200_codes_to_utf8(eng.h, codes.data_ptr(), B, K, T, fr, unicode_offset, codebook_size,
                                             out.data_ptr(), stride, lens, torch.cuda.current_stream().cuda_stream)
        _lib.check(eng.lib, eng.h, rc, "mimi_b200_codes_to_utf8")
    return out, [int(v) for v in lens]


def codes_to_utf8_batch(codes: torch.Tensor, frames: Optional[Sequence[int]] = None, codebook_size: int = CODEBOOK_SIZE,
                        unicode_offset: int = UNICODE_OFFSET, validate: bool = True) -> List[bytes]:
    """codes ``[B,K,T]`` int64 on the GPU -> list of B UTF-8 byte strings (item i uses its first
    ``frames[i]`` frames). One kernel + one device->host copy for the whole batch.

    The kernel's fixed bytes-per-frame layout relies on every code lying in ``[0, codebook_size)``; ``validate`` checks
    that first (``ValueError`` otherwise -- the reference would silently emit characters of a neighbouring codebook or
    raise in ``chr``). Codes that come straight out of the encoder are in range by construction and may skip the check."""
    if codes.dim() != 3:
        raise ValueError("codes must be [batch, num_codebooks, seq_length]")
    if validate and codes.numel():
        lo, hi = int(codes.min()), int(codes.max())
        if lo < 0 or hi >= codebook_size:
            raise ValueError(f"codes must lie in [0, {codebook_size}), got values in [{lo}, {hi}]")
    out, lens = codes_to_utf8_device(codes, frames, codebook_size, unicode_offset)
    if codes.shape[0] == 0:
        return []
    host = out.cpu().numpy()
    return [host[i, : lens[i]].tobytes() for i in range(codes.shape[0])]


def codes_to_chars(codes: Union[List[List[int]], np.ndarray, torch.Tensor], codebook_size: int,
                   copy_before_conversion: bool = True, unicode_offset: int = UNICODE_OFFSET) -> str:
    """REF/emilia-mimi/utils.py:18-37. ``codes`` [K,T] -> str of K*T private-use characters, frame-major.
    The input is never modified (``copy_before_conversion`` is accepted for signature parity)."""
    if isinstance(codes, list):
        codes = np.array(codes)
    if isinstance(codes, np.ndarray):
        if len(codes.shape) != 2:
            raise ValueError("codes must be a 2D array of shape (num_codebooks, seq_length).")
        codes = torch.from_numpy(np.ascontiguousarray(codes).astype(np.int64))
    if len(codes.shape) != 2:
        raise ValueError("codes must be a 2D array of shape (num_codebooks, seq_length).")
    if not codes.is_cuda:
        codes = codes.to(_Engine.get().device)
    return codes_to_utf8_batch(codes[None], None, codebook_size, unicode_offset)[0].decode("utf-8")


def chars_to_codes(chars: str, num_codebooks: int, codebook_size: int, return_tensors: Optional[str] = None,
                   unicode_offset: int = UNICODE_OFFSET):
    """REF/emilia-mimi/utils.py:40-55 (host side; the decode direction is outside the hot path)."""
    cp = np.frombuffer(chars.encode("utf-32-le"), dtype=np.uint32).astype(np.int64)
    codes = cp.reshape(-1, num_codebooks).T.copy()
    codes -= (unicode_offset + np.arange(num_codebooks, dtype=np.int64) * codebook_size)[:, None]
    if return_tensors is None:
        return codes.tolist()
    if return_tensors == "pt":
        return torch.tensor(codes)
    return codes


def audio_to_str(audio_numpy: np.ndarray, mimi_model, device: str = "cuda") -> str:
    """REF/emilia-mimi/utils.py:58-69: un-batched encode, first 8 codebooks, -> unicode string."""
    audio_tensor = torch.tensor(audio_numpy).to(device).unsqueeze(0)
    if len(audio_tensor.shape) == 2:
        audio_tensor = audio_tensor.unsqueeze(1)
    with torch.no_grad():
        audio_codes = mimi_model.encode(audio_tensor)
    codes = audio_codes[0][0][:NUM_CODEBOOKS, :]
    return codes_to_chars(codes, codebook_size=CODEBOOK_SIZE)


def str_to_audio(audio_str: str, mimi_model, device: str = "cuda") -> np.ndarray:
    """REF/emilia-mimi/utils.py:72-81: unicode string -> 8 codebooks -> ``mimi_model.decode`` -> waveform ``[1, 1920*T]``."""
    codes = chars_to_codes(audio_str, num_codebooks=NUM_CODEBOOKS, codebook_size=CODEBOOK_SIZE, return_tensors="pt")
    codes = codes.to(device).unsqueeze(0)
    with torch.no_grad():
        audio_decoded = mimi_model.decode(codes).audio_values[0]
    return audio_decoded.cpu().numpy()


def codes_to_uint16(codes: torch.Tensor) -> torch.Tensor:
    """int64 codes of any shape on the GPU -> uint16 tensor of the same shape: the ``codes.astype(np.uint16)`` storage
    format of REF/yodas2-mimi/process_shard.py:519-523, cast before the device->host copy (2 instead of 8 bytes per code)."""
    if not codes.is_cuda:
        raise _lib.MimiB200Error("codes must live on the CUDA device (no CPU fallback)")
    codes = codes.to(torch.int64).contiguous()
    out = torch.empty(codes.shape, dtype=torch.uint16, device=codes.device)
    eng = _Engine.get(codes.device)
    with eng.lock, torch.cuda.device(codes.device):
        rc = eng.lib.mimi_b200_codes_pack_u16(eng.h, codes.data_ptr(), codes.numel(), out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream)
        _lib.check(eng.lib, eng.h, rc, "mimi_b200_codes_pack_u16")
    return out


def resample_batch(audio: Sequence[np.ndarray], orig_sr: int, target_sr: int = 24000, device=None,
                   pad_to: Optional[int] = None) -> "tuple[torch.Tensor, List[int]]":
    """Resample a list of mono clips on the GPU into one zero-right-padded ``[B,1,N]`` device tensor
    (exactly the ``input_values`` layout) and return it with the per-item output lengths.

    NOT bit-compatible with the reference's ``librosa.resample`` (soxr_hq): see :func:`resample_audio`."""
    eng = _Engine.get(device)
    B = len(audio)
    arrs = [np.ascontiguousarray(np.asarray(a, dtype=np.float32)) for a in audio]
    for a in arrs:
        if a.ndim != 1:
            raise ValueError(f"Expected mono audio but example has shape {a.shape}")
    lens_in = [int(a.shape[0]) for a in arrs]
    lens_out = [int(eng.lib.mimi_b200_resample_out_len(n, orig_sr, target_sr)) for n in lens_in]
    n_in, n_out = max(lens_in + [1]), max(lens_out + [1])
    if pad_to is not None:
        n_out = max(n_out, int(pad_to))
    d_out = torch.empty((B, 1, n_out), dtype=torch.float32, device=eng.device)
    if B == 0:
        return d_out, lens_out
    with eng.lock, torch.cuda.device(eng.device):
        pinned, dev = eng.staging(B * n_in)
        host = pinned[: B * n_in].view(B, n_in)
        # native gather (memcpy threads, GIL released); rows are only read up to their length, no zero fill needed
        src = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
        ln = (C.c_int64 * B)(*lens_in)
        rc = eng.lib.mimi_b200_host_pack(host.data_ptr(), n_in, src, ln, ln, B, 4)
        _lib.check(eng.lib, None, rc, "mimi_b200_host_pack")
        d_in = dev[: B * n_in].view(B, n_in)
        d_in.copy_(host, non_blocking=True)
        eng._staged = torch.cuda.Event()
        eng._staged.record()
        rc = eng.lib.mimi_b200_resample(eng.h, d_in.data_ptr(), n_in, ln, B, int(orig_sr), int(target_sr),
                                        d_out.data_ptr(), n_out, torch.cuda.current_stream().cuda_stream)
        _lib.check(eng.lib, eng.h, rc, "mimi_b200_resample")
    return d_out, lens_out


_warned_resampler = False


def _resampler_backend(backend: Optional[str]) -> str:
    """``None`` -> "librosa" when it is importable (the reference's own resampler: identical tokens), else "b200" with a
    one-time warning that the tokens of resampled audio will not be the reference's."""
    global _warned_resampler
    if backend is not None:
        if backend not in ("librosa", "b200"):
            raise ValueError(f"unknown resampler backend '{backend}'")
        return backend
    import importlib.util
    if importlib.util.find_spec("librosa") is not None:
        return "librosa"
    if not _warned_resampler:
        import warnings
        warnings.warn("resample_audio: librosa is not installed, using the GPU polyphase resampler -- a different filter than "
                      "the reference's soxr_hq, so Mimi codes of the resampled audio will differ from the reference pipeline's "
                      "(pass backend='b200' to choose it explicitly and silence this warning)", RuntimeWarning, stacklevel=3)
        _warned_resampler = True
    return "b200"


def resample_audio(audio: np.ndarray, orig_sr: int, target_sr: int, backend: Optional[str] = None) -> np.ndarray:
    """REF/emilia-mimi/utils.py:84-87: no-op when the rates agree, else band-limited resampling to
    ``ceil(n * target_sr / orig_sr)`` samples (librosa ``fix=True``).

    **Not bit-compatible with the reference.** The reference calls ``librosa.resample``, whose default backend is libsoxr
    "HQ"; that library is not available offline and nothing in the reference pins its output, so parity of this function
    is UNPINNED. ``backend="b200"`` runs the GPU polyphase FIR -- a Kaiser-windowed sinc, 32
    zero crossings, -6 dB at 0.945 x Nyquist -- which is a different and SOFTER filter than soxr_hq (flat to 0.9136 x Nyquist,
    120 dB down from 1.0 x Nyquist): -1.8 dB at soxr_hq's pass-band edge, -24 dB at Nyquist, 120 dB only from 1.08 x Nyquist
    (oracle/resample_oracle.py pins the response). The 24 kHz waveform therefore differs from soxr_hq's in the top sixth of
    the input band, and Mimi codes computed from it differ on the frames where that matters (``bench.py --workload c1`` reports the code agreement between this filter and two other
    high-quality resamplers as a yardstick). ``backend="librosa"`` is the reference's own call (host CPU; needs librosa +
    soxr; ImportError if absent, there is no silent substitute). The default, ``backend=None``, keeps the drop-in
    parity-safe: librosa when it is installed -- as it is wherever the reference scripts run -- and otherwise the GPU kernel
    with a one-time ``RuntimeWarning``. The GPU resampler is thus opt-in (here, ``resample_batch`` and
    ``MimiEncoder.encode_native_rate_batch``) wherever the reference's exact tokens are reachable."""
    if orig_sr == target_sr:
        return audio
    backend = _resampler_backend(backend)
    if backend == "librosa":
        import librosa
        return librosa.resample(audio, orig_sr=orig_sr, target_sr=target_sr)
    out, lens = resample_batch([np.asarray(audio)], orig_sr, target_sr)
    return out[0, 0, : lens[0]].cpu().numpy()
