"""Accuracy experiment for the 3xTF32 accumulation scheme (tc_gemm2): separate vs single accumulator, chunk size."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tokenize_audio_b200 import _lib
lib = _lib.load_library(); h = C.c_void_p()
_lib.check(lib, None, lib.mimi_b200_create(C.byref(h), 0), "create")
lib.mimi_b200_debug_set(h, 3, 2)
for (M, N, K) in [(512, 128, 512), (512, 128, 2048), (512, 128, 8192)]:
    g = torch.Generator().manual_seed(K)
    for dist in ("randn", "pos"):
        a = torch.randn(M, K, generator=g); w = torch.randn(N, K, generator=g) / K ** 0.5
        if dist == "pos":
            a = a.abs(); w = w.abs()
        ad = a.cuda(); wn = np.ascontiguousarray(w.numpy())
        ref = a.double() @ w.double().T
        for single, ck in [(0, 0), (0, 2), (0, 8), (1, 4), (1, 2), (1, 1), (1, 8)]:
            lib.mimi_b200_debug_set(h, 4, single); lib.mimi_b200_debug_set(h, 5, ck)
            out = torch.empty(M, N, device="cuda")
            rc = lib.mimi_b200_debug_tc_gemm(h, ad.data_ptr(), wn.ctypes.data, None, M, N, K, 0, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            _lib.check(lib, h, rc, "gemm")
            d = out.cpu().double() - ref
            print(f"K={K:5d} {dist:5s} single={single} chunk={ck or 4}: rel L2 {float(d.norm()/ref.norm()):.3e}  mean signed rel {float((d/ref.abs().clamp_min(1e-3)).mean()):+.3e}", flush=True)
