"""Time the fused training kernels around the GEMMs in isolation (hot: CUDA graph of 20 back-to-back launches;
cold: one launch after an L2 flush).  Usage (GPU box): python tools/glue_bench.py"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bcnf_b200 import _cabi, train

DEV = torch.device("cuda:0")
L = _cabi.lib()
B, D, H, Cn = 256, 19, 526, 1360
da, dout = 10, 9
HP = 528
stream = lambda: torch.cuda.current_stream().cuda_stream
t = lambda *s: torch.randn(*s, device=DEV)

y, W1, P = t(B, D), t(H, da + Cn) / 30, t(B, HP)
pre, act, d_pre = t(B, HP), t(B, HP), t(B, HP)
Wout, bout = t(2 * dout, H) / 20, t(2 * dout)
ld, ls, ydst, y_out = torch.zeros(B, device=DEV), t(B, dout).tanh(), t(B, dout), t(B, D)
Q = torch.linalg.qr(t(D, D))[0].contiguous()
scale, bias, xsave = t(D).abs() + 0.5, t(D), t(B, D)
gs, gb = torch.zeros(D, device=DEV), torch.zeros(D, device=DEV)
dz, dz_out, dld, d_o = t(B, D), t(B, D), t(B), t(B, 2 * dout)
img = train._Img(DEV, B, H)


def ops(arr):
    arr[0].type, arr[0].p0 = 0, Q.data_ptr()
    arr[1].type, arr[1].p0, arr[1].p1, arr[1].save = 1, scale.data_ptr(), bias.data_ptr(), xsave.data_ptr()
    arr[1].g0, arr[1].g1 = gs.data_ptr(), gb.data_ptr()
    return 2


def k_pre():
    a = _cabi.TrainPreArgs()
    a.y, a.y_pitch, a.B, a.D, a.src0, a.din = y.data_ptr(), D, B, D, 0, da
    a.W1, a.w1_pitch, a.P, a.p_pitch, a.H = W1.data_ptr(), W1.stride(0), P.data_ptr(), HP, H
    a.pre, a.act, a.pitch = pre.data_ptr(), act.data_ptr(), HP
    a.seed, a.layer_uid, a.p_drop = 5, 1, 0.4
    a.act_img, a.img_plane, a.img_rpad = img.ptr, img.plane, img.rpad
    return lambda: _cabi.check(L.bcnf_train_pre(C.byref(a), 0, stream()), "pre")


def k_post():
    a = _cabi.TrainPostArgs()
    a.a, a.a_pitch, a.Wout, a.bout = act.data_ptr(), HP, Wout.data_ptr(), bout.data_ptr()
    a.B, a.D, a.H, a.dst0, a.dout = B, D, H, da, dout
    a.y_in, a.y_out, a.ld, a.ls_save, a.ydst_save = y.data_ptr(), y_out.data_ptr(), ld.data_ptr(), ls.data_ptr(), ydst.data_ptr()
    a.n_ops = ops(a.ops)
    return lambda: _cabi.check(L.bcnf_train_post(C.byref(a), 0, stream()), "post")


def k_post_bwd():
    a = _cabi.TrainPostBwdArgs()
    a.dz_in, a.dz_out, a.dld = dz.data_ptr(), dz_out.data_ptr(), dld.data_ptr()
    a.B, a.D, a.H, a.dst0, a.dout = B, D, H, da, dout
    a.ls_save, a.ydst_save, a.Wout = ls.data_ptr(), ydst.data_ptr(), Wout.data_ptr()
    a.pre, a.pitch, a.d_o, a.d_pre = pre.data_ptr(), HP, d_o.data_ptr(), d_pre.data_ptr()
    a.seed, a.layer_uid, a.p_drop = 5, 4, 0.4
    a.n_ops = ops(a.ops)
    a.dpre_img, a.img_plane, a.img_rpad = img.ptr, img.plane, img.rpad
    return lambda: _cabi.check(L.bcnf_train_post_bwd(C.byref(a), 0, stream()), "post_bwd")


def k_pre_bwd():
    a = _cabi.TrainPreBwdArgs()
    a.d_pre, a.pitch, a.W1, a.w1_pitch = d_pre.data_ptr(), HP, W1.data_ptr(), W1.stride(0)
    a.B, a.D, a.H, a.src0, a.din, a.dz = B, D, H, 0, da, dz_out.data_ptr()
    return lambda: _cabi.check(L.bcnf_train_pre_bwd(C.byref(a), 0, stream()), "pre_bwd")


def k_colsum():
    out = torch.zeros(H, device=DEV)
    return lambda: train._colsum(d_pre, out, cols=H)


def k_pack():
    W = t(H, H)
    f, b = train._Img(DEV, H, H), train._Img(DEV, H, H)
    descs = []
    for _ in range(5):
        descs += [(W, 0, H, 1, H, H, f), (W, 0, 1, H, H, H, b)]
    return lambda: train._pack_images(descs, DEV)


flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
for name, mk in [("pre", k_pre), ("post", k_post), ("post_bwd", k_post_bwd), ("pre_bwd", k_pre_bwd), ("colsum", k_colsum),
                 ("img_pack x10", k_pack)]:
    fn = mk()
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    hot = e0.elapsed_time(e1) / 100 * 1e3
    cold = []
    for _ in range(5):
        flush.fill_(1); torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        cold.append(e0.elapsed_time(e1) * 1e3)
    print(f"{name:14s} hot {hot:7.2f} us   cold (after L2 flush) {min(cold):7.2f} us")
