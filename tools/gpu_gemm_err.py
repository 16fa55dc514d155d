#!/usr/bin/env python3
"""Relative error of the debug GEMM hook in a given kernel generation: python tools/gpu_gemm_err.py MODE M N K ..."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tokenize_audio_b200 import _lib
lib = _lib.load_library(); h = C.c_void_p()
_lib.check(lib, None, lib.mimi_b200_create(C.byref(h), 0), "create")
mode = int(sys.argv[1]); lib.mimi_b200_debug_set(h, 3, mode)
v = [int(a) for a in sys.argv[2:]]
for M, N, K in zip(v[0::3], v[1::3], v[2::3]):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5
    ad = a.cuda(); wn = np.ascontiguousarray(w.numpy()); out = torch.zeros(M, N, device="cuda")
    rc = lib.mimi_b200_debug_tc_gemm(h, ad.data_ptr(), wn.ctypes.data, None, M, N, K, 0, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(lib, h, rc, "gemm")
    ref = a.double() @ w.double().T
    o = out.cpu().double()
    hi = a.double().clone()
    # what the result would be with the A_lo * W_hi term missing entirely
    ah = torch.from_numpy(((a.numpy().view(np.uint32) + 0x1000) & 0xFFFFE000).view(np.float32)).double()
    ref_nolo = ah @ w.double().T
    print(f"mode {mode} M={M} N={N} K={K}: rel err {float((o - ref).norm() / ref.norm()):.3e}   (no A_lo term would be {float((ref_nolo - ref).norm() / ref.norm()):.3e})", flush=True)
