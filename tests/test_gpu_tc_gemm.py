"""Unit test of the CTA-pair tcgen05 GEMM kernel in both split-precision generations (through the C-ABI debug hook) against
float64 matmul. Tolerance: relative L2 error <= 3e-6 (fp32-equivalent), far inside the 3e-5 the codes tolerate."""
import ctypes as C

import numpy as np
import pytest
import torch

from tokenize_audio_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    lib = _lib.load_library()
    h = C.c_void_p()
    _lib.check(lib, None, lib.mimi_b200_create(C.byref(h), 0), "create")
    yield lib, h
    lib.mimi_b200_destroy(h)


@pytest.mark.parametrize("mode", [9, 7])
@pytest.mark.parametrize("M,N,K,act,bias", [
    (128, 128, 32, 0, False), (300, 128, 512, 0, True), (77, 64, 384, 0, True), (1000, 256, 1280, 0, False),
    (60, 2048, 512, 1, False), (130, 512, 2048, 0, True), (257, 1024, 8192, 0, True), (64, 1536, 512, 0, False),
    (40000, 128, 64, 0, True), (25000, 64, 96, 0, False),
])
def test_tc_gemm_matches_float64(engine, M, N, K, act, bias, mode):
    lib, h = engine
    _lib.check(lib, h, lib.mimi_b200_debug_set(h, 3, mode), "debug_set")
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g) * 2.0
    w = torch.randn(N, K, generator=g) / K ** 0.5
    bvec = torch.randn(N, generator=g) if bias else None
    ad = a.cuda()
    bd = bvec.cuda() if bias else None
    out = torch.empty(M, N, device="cuda")
    wn = np.ascontiguousarray(w.numpy())
    rc = lib.mimi_b200_debug_tc_gemm(h, ad.data_ptr(), wn.ctypes.data, bd.data_ptr() if bias else None, M, N, K, act,
                                     out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(lib, h, rc, "debug_tc_gemm")
    ref = a.double() @ w.double().T
    if bias:
        ref = ref + bvec.double()
    if act:
        ref = torch.nn.functional.gelu(ref)
    err = (out.cpu().double() - ref).norm() / ref.norm()
    # mode 7 keeps lo and the W_hi of the A_lo * W_hi term in bf16; mode 9: fp16 hi/lo split of both operands (22 bits each)
    tol = 3e-6
    assert err <= tol, f"relative error {err:.2e}"
