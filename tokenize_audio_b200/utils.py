"""Drop-in counterparts of the reference's ``utils.py`` (REF/emilia-mimi/utils.py, six identical copies)
with the array work done by libmimi_b200.so kernels: ``codes_to_chars``, ``chars_to_codes``,
``audio_to_str``, ``resample_audio``. Same names, argument meaning and error behaviour.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

UNICODE_OFFSET: int = 0xE000      # REF/emilia-mimi/utils.py:13-15
NUM_CODEBOOKS: int = 8
CODEBOOK_SIZE: int = 2048


class _Engine:
    """Weight-less engine handle for the resampler / UTF-8 kernels (one per device)."""

    _cache = {}

    def __init__(self, device: torch.device):
        self.lib = _lib.load_library()
        self.device = device
        h = C.c_void_p()
        _lib.check(self.lib, None, self.lib.mimi_b200_create(C.byref(h), device.index), "mimi_b200_create")
        self.h = h

    @classmethod
    def get(cls, device=None) -> "_Engine":
        if not torch.cuda.is_available():
            raise _lib.MimiB200Error("a CUDA device (B200) is required; there is no CPU fallback")
        dev = torch.device(device if device is not None else "cuda")
        dev = torch.device("cuda", dev.index if dev.index is not None else torch.cuda.current_device())
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


def codes_to_utf8_batch(codes: torch.Tensor, frames: Optional[Sequence[int]] = None, codebook_size: int = CODEBOOK_SIZE,
                        unicode_offset: int = UNICODE_OFFSET) -> List[bytes]:
    """codes ``[B,K,T]`` int64 on the GPU -> list of B UTF-8 byte strings (item i uses its first
    ``frames[i]`` frames). One kernel + one device->host copy for the whole batch."""
    if codes.dim() != 3:
        raise ValueError("codes must be [batch, num_codebooks, seq_length]")
    B, K, T = codes.shape
    validate_unicode_offset(unicode_offset, K, codebook_size)
    eng = _Engine.get(codes.device)
    bpf = eng.lib.mimi_b200_utf8_bytes_per_frame(K, unicode_offset, codebook_size)
    if bpf < 0:
        raise ValueError("unicode offset / codebook size not representable as fixed-width UTF-8 per codebook")
    if B == 0:
        return []
    codes = codes.to(torch.int64).contiguous()
    stride = max(int(T * bpf), 1)
    out = torch.empty((B, stride), dtype=torch.uint8, device=codes.device)
    fr = (C.c_int64 * B)(*[int(f) for f in frames]) if frames is not None else None
    lens = (C.c_int64 * B)()
    with torch.cuda.device(codes.device):
        rc = eng.lib.mimi_b200_codes_to_utf8(eng.h, codes.data_ptr(), B, K, T, fr, unicode_offset, codebook_size,
                                             out.data_ptr(), stride, lens, torch.cuda.current_stream().cuda_stream)
    _lib.check(eng.lib, eng.h, rc, "mimi_b200_codes_to_utf8")
    host = out.cpu().numpy()
    return [host[i, : lens[i]].tobytes() for i in range(B)]


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


def resample_batch(audio: Sequence[np.ndarray], orig_sr: int, target_sr: int = 24000, device=None,
                   pad_to: Optional[int] = None) -> "tuple[torch.Tensor, List[int]]":
    """Resample a list of mono clips on the GPU into one zero-right-padded ``[B,1,N]`` device tensor
    (exactly the ``input_values`` layout) and return it with the per-item output lengths."""
    eng = _Engine.get(device)
    B = len(audio)
    lens_in = [int(len(a)) for a in audio]
    lens_out = [int(eng.lib.mimi_b200_resample_out_len(n, orig_sr, target_sr)) for n in lens_in]
    n_in, n_out = max(lens_in + [1]), max(lens_out + [1])
    if pad_to is not None:
        n_out = max(n_out, int(pad_to))
    host = torch.zeros((B, n_in), dtype=torch.float32).pin_memory()
    for i, a in enumerate(audio):
        host[i, : lens_in[i]] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    d_in = host.to(eng.device, non_blocking=True)
    d_out = torch.empty((B, 1, n_out), dtype=torch.float32, device=eng.device)
    hl = (C.c_int64 * B)(*lens_in)
    with torch.cuda.device(eng.device):
        rc = eng.lib.mimi_b200_resample(eng.h, d_in.data_ptr(), n_in, hl, B, int(orig_sr), int(target_sr),
                                        d_out.data_ptr(), n_out, torch.cuda.current_stream().cuda_stream)
    _lib.check(eng.lib, eng.h, rc, "mimi_b200_resample")
    return d_out, lens_out


def resample_audio(audio: np.ndarray, orig_sr: int, target_sr: int) -> np.ndarray:
    """REF/emilia-mimi/utils.py:84-87: no-op when the rates agree, else band-limited resampling to
    ``ceil(n * target_sr / orig_sr)`` samples (librosa ``fix=True``), here by the GPU polyphase FIR."""
    if orig_sr == target_sr:
        return audio
    out, lens = resample_batch([np.asarray(audio)], orig_sr, target_sr)
    return out[0, 0, : lens[0]].cpu().numpy()
