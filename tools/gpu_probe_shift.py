"""Hardware probe: does a SWIZZLE_128B K-major UMMA operand descriptor work with a start address shifted by
whole 128-byte rows? Prints the relative error of out[m] = A[m + shift] @ W^T (single-pass TF32, so ~1e-3
is "correct") for shift 0..8 with the descriptor base-offset field left 0 / set to (start >> 7) & 7."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from tokenize_audio_b200 import _lib

lib = _lib.load_library()
h = C.c_void_p()
_lib.check(lib, None, lib.mimi_b200_create(C.byref(h), 0), "create")
K = 96
g = torch.Generator().manual_seed(1)
a = torch.randn(136, K, generator=g)
w = torch.randn(64, K, generator=g)
ad = a.cuda()
wn = np.ascontiguousarray(w.numpy())
for base_mode in (0, 1):
    for shift in range(9):
        out = torch.zeros(128, 64, device="cuda")
        rc = lib.mimi_b200_debug_shift_probe(h, ad.data_ptr(), wn.ctypes.data, K, shift, base_mode, out.data_ptr(),
                                             torch.cuda.current_stream().cuda_stream)
        _lib.check(lib, h, rc, "shift_probe")
        ref = a[shift:shift + 128].double() @ w.double().T
        err = float((out.cpu().double() - ref).norm() / ref.norm())
        # which shift does the result actually correspond to?
        best = min(range(9), key=lambda s: float((out.cpu().double() - a[s:s + 128].double() @ w.double().T).norm()))
        print(f"base_mode={base_mode} shift={shift}: rel err {err:.3e}  (closest to shift {best})", flush=True)
lib.mimi_b200_destroy(h)
