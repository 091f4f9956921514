"""Pipeline stamps (clock64, CTA (0,0,0)) of the tensor-core training GEMM.  Usage: python tools/tc_gemm_trace.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bcnf_b200 import _cabi

DEV = torch.device("cuda:0")
L = _cabi.lib()


def trace(kind, M, N, K, bn):
    X = torch.randn(M, K, device=DEV)
    if kind == "fwd":
        W = torch.randn(N, K, device=DEV); a, astr, b, bstr = X, (K, 1), W, (1, K)
    elif kind == "dx":
        W = torch.randn(K, N, device=DEV); a, astr, b, bstr = X, (K, 1), W, (N, 1)
    else:
        D = torch.randn(K, M, device=DEV); Xx = torch.randn(K, N, device=DEV); a, astr, b, bstr = D, (1, M), Xx, (N, 1)
    Cm = torch.empty(M, N, device=DEV)
    g = _cabi.GemmArgs()
    g.A, g.B, g.C = a.data_ptr(), b.data_ptr(), Cm.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.as0, g.as1 = astr
    g.bs0, g.bs1 = bstr
    g.cs0 = N
    g.split_k = 1
    old = L.bcnf_train_set_gemm_mode(2 | (bn << 4))
    out = (C.c_int64 * 64)()
    for _ in range(3):   # warm: the last run is the one reported
        _cabi.check(L.bcnf_train_gemm_trace(C.byref(g), 0, torch.cuda.current_stream().cuda_stream, out), "trace")
    L.bcnf_train_set_gemm_mode(old)
    t = list(out)
    t0 = t[0]
    rel = lambda v: (v - t0) if v else None
    print(f"--- {kind} M={M} N={N} K={K} BN={bn}: cycles since CTA start")
    print(f"setup done {rel(t[1])}; loaders done {rel(t[2])}; acc_full {rel(t[3])}; epilogue done {rel(t[4])}; exit {rel(t[5])}")
    for it in range(min(8, (K + 63) // 64)):
        print(f"  chunk {it}: loads issued {rel(t[16 + it])}  stored {rel(t[24 + it])}  handed {rel(t[32 + it])}  mma saw full {rel(t[8 + it])}")


if __name__ == "__main__":
    trace("fwd", 256, 526, 526, 32)
    trace("fwd", 256, 526, 526, 64)
    trace("dw", 526, 526, 256, 32)
    trace("dx", 256, 526, 526, 32)
    trace("fwd", 256, 512, 512, 32)
