"""Check and time the tensor-core training GEMM (csrc/train_tc.cuh) against torch fp64 / the FMA kernel.
Usage (GPU box): python tools/tc_gemm_check.py"""
import itertools
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bcnf_b200 import _cabi, train

DEV = torch.device("cuda:0")
L = _cabi.lib()


def run(kind, M, N, K, mode, split_k=0, iters=0):
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    if kind == "fwd":      # C = X W^T : A (M,K) k-fast ; B(r,j) = W[j,r]
        X = torch.randn(M, K, generator=g).to(DEV); W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
        a, astr, b, bstr = X, (K, 1), W, (1, K)
        ref = X.double() @ W.double().t()
    elif kind == "dx":     # C = D W : A (M,K) ; B(r,j) = W[r,j]  (W is (K,N))
        X = torch.randn(M, K, generator=g).to(DEV); W = (torch.randn(K, N, generator=g) / K ** 0.5).to(DEV)
        a, astr, b, bstr = X, (K, 1), W, (N, 1)
        ref = X.double() @ W.double()
    else:                  # dw: C = D^T X : A(i,r) = D[r,i] (D is (K,M)) ; B(r,j) = X[r,j] (X is (K,N))
        D = torch.randn(K, M, generator=g).to(DEV); X = torch.randn(K, N, generator=g).to(DEV)
        a, astr, b, bstr = D, (1, M), X, (N, 1)
        ref = D.double().t() @ X.double()
    C = torch.empty(M, N, device=DEV)
    old = L.bcnf_train_set_gemm_mode(mode)
    try:
        train._gemm(a, astr, b, bstr, C, M, N, K, split_k=split_k)
        torch.cuda.synchronize()
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
        ms = None
        if iters:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            graph = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                train._gemm(a, astr, b, bstr, C, M, N, K, split_k=split_k)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(graph):
                for _ in range(20):
                    train._gemm(a, astr, b, bstr, C, M, N, K, split_k=split_k)
            graph.replay(); torch.cuda.synchronize()
            e0.record()
            for _ in range(iters):
                graph.replay()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / (iters * 20)
    finally:
        L.bcnf_train_set_gemm_mode(old)
    return err, ms


def run_img(M, N, K, bn, iters=5):
    """TMA-fed kernel on operand images (forward orientation; the data-gradient GEMM is the same kernel)."""
    g = torch.Generator().manual_seed(1)
    X = torch.randn(M, K, generator=g).to(DEV); W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    xi, wi = train._Img(DEV, M, K), train._Img(DEV, N, K)
    train._pack_images([(X, 0, K, 1, M, K, xi), (W, 0, K, 1, N, K, wi)], DEV)
    C = torch.empty(M, N, device=DEV)
    old = L.bcnf_train_set_gemm_mode(bn << 4)
    call = lambda: train._gemm(None, None, None, None, C, M, N, K, split_k=1, a_img=xi, b_img=wi)
    try:
        call(); torch.cuda.synchronize()
        ref = X.double() @ W.double().t()
        err = ((C.double() - ref).abs().max() / ref.abs().max()).item()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                call()
        graph.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            graph.replay()
        e1.record(); torch.cuda.synchronize()
        return err, e0.elapsed_time(e1) / (iters * 20)
    finally:
        L.bcnf_train_set_gemm_mode(old)


if __name__ == "__main__":
    DEV = torch.device("cuda:0")
    if "--ncu" in sys.argv:      # one shape: for ncu --set full -k regex:train_tc2 -s 2 -c 1
        err, ms = run_img(256, 526, 526, 32, iters=1)
        print(f"hidden-layer GEMM at batch 256, bn32: rel err {err:.2e}, {ms * 1e3:.1f} us")
        sys.exit(0)
    print("== TMA-fed image kernel: rel. error vs fp64, us per GEMM (CUDA graph of 20 back-to-back launches) ==")
    for M, N, K in [(256, 526, 526), (256, 526, 1360), (256, 1360, 526), (4096, 526, 526), (32768, 526, 526)]:
        row = []
        for bn in (32, 64, 128):
            err, ms = run_img(M, N, K, bn)
            row.append(f"bn{bn}: {ms * 1e3:7.1f} us ({2 * M * N * K / ms / 1e9:6.1f} TFLOP/s, err {err:.1e})")
        print(f"M={M:5d} N={N:4d} K={K:4d}: " + " | ".join(row))
    if "--img-only" in sys.argv:
        sys.exit(0)
    print("== correctness (rel. max error vs fp64) ==")
    for kind, (M, N, K) in itertools.product(["fwd", "dx", "dw"], [(77, 53, 90), (256, 526, 526), (130, 40, 1370), (526, 1370, 256), (300, 128, 64)]):
        for mode, name in [(1, "fma"), (2, "tc64"), (2 | (32 << 4), "tc32"), (2 | (128 << 4), "tc128")]:
            for sk in ([1] if mode == 1 else [1, 0]):
                err, _ = run(kind, M, N, K, mode, split_k=sk)
                print(f"{kind:3s} M={M:4d} N={N:4d} K={K:4d} {name:5s} split={'no' if sk == 1 else 'auto'}: err {err:.2e}")
    print("== time per GEMM (us), CUDA graph of 20 back-to-back launches ==")
    shapes = [("fwd", 256, 526, 526), ("dx", 256, 526, 526), ("dw", 526, 526, 256), ("fwd", 256, 526, 1370), ("dx", 256, 1370, 526),
              ("dw", 526, 1370, 256), ("fwd", 4096, 526, 526), ("dx", 4096, 526, 526), ("dw", 526, 526, 4096)]
    for kind, M, N, K in shapes:
        row = []
        for mode, name in [(1, "fma"), (2 | (32 << 4), "tc32"), (2 | (64 << 4), "tc64"), (2 | (128 << 4), "tc128")]:
            for sk in ([1] if mode == 1 else [1, 0]):
                err, ms = run(kind, M, N, K, mode, split_k=sk, iters=5)
                row.append(f"{name}{'' if sk == 1 else '+splitK'} {ms * 1e3:7.1f}")
        print(f"{kind:3s} M={M:4d} N={N:4d} K={K:4d}: " + " | ".join(row), f" ({2 * M * N * K / 1e6:.0f} MFLOP)")
