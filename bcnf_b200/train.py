"""Training step of the coupling stack (reference ``Trainer._train_batch``, src/bcnf/train/trainer.py:244-277).

``CondRealNVP_v2.forward`` in training mode routes here: one ``torch.autograd.Function`` for the whole stack.
The conditioner's Linear -> GELU -> Dropout chains -- all of the FLOPs -- run forward and backward on the
tensor-core GEMMs of ``csrc/train_tc.cuh`` through the C ABI (``bcnf_train_gemm`` on bf16 hi / lo operand
images): bias + GELU + dropout in the forward epilogue, gelu' * mask in the data-gradient epilogue, weight
gradients written straight into tensors shaped like the parameters (or into the flat gradient buffer of
``_GradSink``).  Dropout masks are a counter-based hash of (seed, layer, row, column), regenerated in backward,
never stored.  Everything between the GEMMs of a coupling block is one fused kernel per direction
(``csrc/train_glue.cuh``).

Gradients are per-parameter tensors (views of one flat buffer when the Trainer owns the step), so ``torch.optim``
and ``torch.nn.parallel.DistributedDataParallel`` work unchanged on top; ``FlatAdam`` and ``fused_nll`` are the
fused forms of the optimizer step and the loss.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Any, Sequence

import torch

from . import _cabi

__all__ = ["stack_forward_train", "dropout_mask", "Trainer", "FlatAdam", "fused_nll"]


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


_WS: dict[int, tuple[torch.Tensor, torch.Tensor]] = {}
_WS_FLOATS = 148 * 128 * 128        # one 128 x 128 fp32 tile per SM: split-K only runs when the tiles alone leave SMs idle
_WS_COUNTERS = 1024


def _workspace(dev: torch.device) -> tuple[torch.Tensor, torch.Tensor]:
    """Split-K scratch of the tensor-core GEMM (zero on entry, left zero by every launch); one per device, shared by
    all GEMMs of that device -- they are ordered on one stream (the step runs on the current stream)."""
    key = dev.index or 0
    if key not in _WS:
        _WS[key] = (torch.zeros(_WS_FLOATS, device=dev), torch.zeros(_WS_COUNTERS, dtype=torch.int32, device=dev))
    return _WS[key]


class _Img:
    """Operand image of a (rows x k) matrix for the TMA-fed tensor-core GEMM (include/bcnf_b200.h): bf16 hi and lo
    planes, [ceil(k/64)][rows padded to 128][128 bytes] each; zero outside the valid extent (allocated zeroed, and
    producers only ever write inside rows x tiles)."""

    def __init__(self, dev: torch.device, rows: int, k: int, align: int = 128) -> None:
        self.rows, self.k = rows, k
        self.rpad = (rows + align - 1) // align * align          # the CTA-pair GEMM wants multiples of 256 rows
        self.chunks = (k + 127) // 128 * 2                       # whole 128-column pairs: a weight-gradient tile reads two chunks
        self.plane = self.chunks * self.rpad * 128
        self.buf = torch.zeros(2 * self.plane, dtype=torch.uint8, device=dev)

    @property
    def ptr(self) -> int:
        return self.buf.data_ptr()


class _PoolLease:
    """A pool of images on loan to one forward / backward pair.  Pools of a given (device, shapes) are recycled through a
    free list -- their padding stays zero, producers only ever write valid data or zeros -- but never shared by two
    calls whose saved state is alive at the same time: the lease returns the pool when the autograd context that
    holds it is released."""
    _free: dict[tuple, list[list[_Img]]] = {}

    def __init__(self, dev: torch.device, shapes: list[tuple[int, int]]) -> None:
        self.key = (dev.index or 0, tuple(shapes))
        free = _PoolLease._free.setdefault(self.key, [])
        self.images = free.pop() if free else [_Img(dev, rows, k) for rows, k in shapes]

    def __del__(self) -> None:
        try:
            _PoolLease._free.setdefault(self.key, []).append(self.images)
        except Exception:      # interpreter shutdown
            pass


import collections

_IMGS: "collections.OrderedDict[tuple, _Img]" = collections.OrderedDict()
_IMG_OWNERS: dict[tuple, Any] = {}      # key -> weakref of the tensor the image belongs to (entry dies with it)
_SCRATCH_CAP = 256                      # ownerless (row-scratch) images kept; least recently used ones are dropped


def _img(dev: torch.device, tag: Any, rows: int, k: int, align: int = 128, owner: torch.Tensor | None = None) -> _Img:
    """Cached scratch image: (device, tag, shape) -> buffer.  Scratch images are reused in stream order.

    ``owner``: the parameter an image is made from.  The entry is then dropped when that tensor is garbage-collected
    (a weakref callback), so that building and discarding models -- k-fold runs, hyper-parameter searches -- does not
    pin two bf16 hi/lo copies of every weight matrix for the life of the process.
    """
    key = (dev.index or 0, tag, rows, k, align)
    im = _IMGS.get(key)
    if im is not None:
        _IMGS.move_to_end(key)
    else:
        im = _IMGS[key] = _Img(dev, rows, k, align)
        if owner is None and len(_IMGS) - len(_IMG_OWNERS) > _SCRATCH_CAP:
            # a new batch / chunk size every call must not pin device memory forever: drop the scratch image that has
            # not been used for longest (its memory returns to the caching allocator, stream-ordered like any tensor)
            for old in _IMGS:
                if old not in _IMG_OWNERS:
                    del _IMGS[old]
                    break
        if owner is not None:
            import weakref

            def _drop(_ref, key=key):
                _IMGS.pop(key, None)
                _IMG_OWNERS.pop(key, None)
            try:
                _IMG_OWNERS[key] = weakref.ref(owner, _drop)
            except TypeError:
                pass
    return im


def clear_caches() -> None:
    """Release every cached operand image and image pool of this process (``Trainer.close()`` calls it)."""
    _IMGS.clear()
    _IMG_OWNERS.clear()
    _PoolLease._free.clear()
    try:
        from . import feature_tc
        clear = getattr(feature_tc, "clear_caches", None)
        if clear is not None:
            clear()
    except Exception:      # interpreter shutdown
        pass


def _pack_images(descs: list[tuple[torch.Tensor, int, int, int, int, int, _Img]], dev: torch.device) -> None:
    """descs: (tensor, element offset, s_row, s_k, rows, k, image).  One launch per 64 descriptors."""
    arr = (_cabi.ImgPackDesc * len(descs))()
    for d, (t, off, s_row, s_k, rows, k, im) in zip(arr, descs):
        d.src, d.s_row, d.s_k, d.rows, d.k = t.data_ptr() + 4 * off, s_row, s_k, rows, k
        d.dst, d.plane, d.rpad, d.chunks = im.ptr, im.plane, im.rpad, im.chunks
    _cabi.check(_cabi.lib().bcnf_img_pack(arr, len(descs), dev.index or 0, _stream(dev)), "bcnf_img_pack")


def _gemm(A, a_strides, B, b_strides, Cm, M, N, K, *, beta=0.0, epi=_cabi.EPI_NONE, bias=None, save=None,
          saved=None, seed=0, uid=0, p=0.0, seed_ptr=None, colsum=None, split_k=0, c_stride=None,
          a_img: _Img | None = None, b_img: _Img | None = None, c_img: _Img | None = None, mn: bool = False) -> None:
    g = _cabi.GemmArgs()
    if split_k != 1:
        ws, counters = _workspace(Cm.device)
        g.ws, g.counters, g.ws_floats, g.n_counters = ws.data_ptr(), counters.data_ptr(), ws.numel(), counters.numel()
    g.split_k = split_k
    g.colsum = colsum.data_ptr() if colsum is not None else None
    if a_img is not None and b_img is not None:
        g.a_img, g.a_plane, g.a_rpad = a_img.ptr, a_img.plane, a_img.rpad
        g.b_img, g.b_plane, g.b_rpad = b_img.ptr, b_img.plane, b_img.rpad
        g.img_mn = 1 if mn else 0
    else:
        g.A, g.B = A.data_ptr(), B.data_ptr()
        g.as0, g.as1 = a_strides
        g.bs0, g.bs1 = b_strides
    if c_img is not None:
        g.c_img, g.c_plane, g.c_rpad = c_img.ptr, c_img.plane, c_img.rpad
    g.C = Cm.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.cs0 = Cm.stride(0) if c_stride is None else c_stride
    g.beta, g.epilogue = beta, epi
    g.bias = bias.data_ptr() if bias is not None else None
    g.save = save.data_ptr() if save is not None else None
    g.saved = saved.data_ptr() if saved is not None else None
    g.seed, g.layer_uid, g.p_drop = seed, uid, p
    g.seed_ptr = seed_ptr
    dev = Cm.device
    _cabi.check(_cabi.lib().bcnf_train_gemm(C.byref(g), dev.index or 0, _stream(dev)), "bcnf_train_gemm")


def _colsum(X: torch.Tensor, out: torch.Tensor, cols: int | None = None) -> None:
    dev = X.device
    _cabi.check(_cabi.lib().bcnf_train_colsum(X.data_ptr(), X.shape[0], X.shape[1] if cols is None else cols, X.stride(0),
                                              out.data_ptr(), 0.0,
                                              dev.index or 0, _stream(dev)), "bcnf_train_colsum")


def dropout_mask(rows: int, cols: int, seed: int, uid: int, p: float, device: Any,
                 seed_word: torch.Tensor | None = None) -> torch.Tensor:
    """The multiplicative mask (0 or 1/(1-p)) the fused epilogues apply for (seed, uid) -- for tests."""
    dev = torch.device(device)
    out = torch.empty(rows, cols, device=dev)
    sp = seed_word.data_ptr() if seed_word is not None else None
    _cabi.check(_cabi.lib().bcnf_train_dropout_mask(out.data_ptr(), rows, cols, seed, uid, p, sp, dev.index or 0, _stream(dev)),
                "bcnf_train_dropout_mask")
    return out


def layer_uid(layer_index: int, net: int, lin: int) -> int:
    """Identifier of one dropout site: (index in model.layers, nn_a=0 / nn_b=1, hidden Linear index)."""
    return (layer_index * 2 + net) * 16 + lin


class _Spec:
    """Static description of the stack handed to the autograd function (not a tensor)."""

    def __init__(self, kinds: list[str], n_lin: int, two_way: bool, size: int, n_conditions: int, p_drop: float,
                 seed: int, seed_word: torch.Tensor | None = None) -> None:
        self.kinds, self.n_lin, self.two_way = kinds, n_lin, two_way
        self.size, self.n_conditions, self.p_drop, self.seed = size, n_conditions, p_drop, seed
        self.seed_word = seed_word          # device int64 word XORed into the seed by the kernels (CUDA-graph replays)
        self.seed_ptr = seed_word.data_ptr() if seed_word is not None else None


# Work off the dependency chain (condition projections, weight / bias gradients, d h) runs on side streams.  A
# weight-gradient GEMM at batch 256 fills ~45 SMs and a chain GEMM needs ~34 free ones, so two side streams leave the
# chain room.  BCNF_TRAIN_SIDE_STREAMS=0 puts everything on the caller's stream (isolated kernel timings).
_N_SIDE = int(os.environ.get("BCNF_TRAIN_SIDE_STREAMS", "2"))
_SIDE: dict[int, list[Any]] = {}
# Operand images of a network's parameters are packed this many networks ahead of the chain (forward orientation in the
# forward pass, data-gradient orientation in the backward pass), so that the GEMMs find them in L2: packed all at once
# at the start of the step (0 = the previous behaviour) the 260 MB of images evict each other before they are read.
_PACK_AHEAD = int(os.environ.get("BCNF_TRAIN_PACK_AHEAD", "2"))
_PACK: dict[int, Any] = {}
# Weight / bias gradients of the Transformer encoder's Linears on the side streams instead of the backward chain
# (feature_network.OffChain; BCNF_TRAIN_ENC_OFF_CHAIN=0: plain nn.Linear autograd)
_ENC_OFF_CHAIN = os.environ.get("BCNF_TRAIN_ENC_OFF_CHAIN", "1") != "0"
# FlatAdam applied bucket by bucket behind the gradient all-reduce, underneath the backward (BCNF_TRAIN_EARLY_ADAM=0: one
# launch at the end of the step)
_EARLY_ADAM = os.environ.get("BCNF_TRAIN_EARLY_ADAM", "1") != "0"


def _pack_stream(dev: torch.device) -> torch.cuda.Stream:
    key = dev.index or 0
    if key not in _PACK:
        _PACK[key] = torch.cuda.Stream(device=dev)
    return _PACK[key]


def _side_streams(dev: torch.device) -> list[Any]:
    if _N_SIDE <= 0:
        return [torch.cuda.current_stream(dev)]
    key = dev.index or 0
    if key not in _SIDE:
        _SIDE[key] = [torch.cuda.Stream(device=dev) for _ in range(_N_SIDE)]
    return _SIDE[key]


def _side_stream(dev: torch.device) -> torch.cuda.Stream:
    return _side_streams(dev)[0]


class _Unit:
    """One conditioner network + the half it transforms + the ActNorm / mixing layers that follow it."""

    def __init__(self, li: int, net: int, w0: int, src0: int, din: int, dst0: int, dout: int) -> None:
        self.li, self.net, self.w0 = li, net, w0          # w0: index of the network's first weight in the parameter list
        self.src0, self.din, self.dst0, self.dout = src0, din, dst0, dout
        self.ops: list[tuple[int, int]] = []              # (glue type, index of its first parameter)


def _plan(spec: _Spec) -> tuple[list[tuple[int, int]], list[_Unit]]:
    """Split model.layers into the glue ops before the first coupling and one _Unit per conditioner network."""
    D = spec.size
    da = (D + 1) // 2
    lead: list[tuple[int, int]] = []
    units: list[_Unit] = []
    o = 0
    for li, kind in enumerate(spec.kinds):
        if kind == "actnorm":
            (units[-1].ops if units else lead).append((_cabi.GLUE_ACTNORM, o))
            o += 2
        elif kind == "ortho":
            (units[-1].ops if units else lead).append((_cabi.GLUE_ORTHO, o))
            o += 1
        else:
            # nn_a reads y_a and transforms y_b; nn_b reads z_b and transforms y_a (cnf.py:178-184)
            units.append(_Unit(li, 0, o, 0, da, da, D - da))
            o += 2 * spec.n_lin
            if spec.two_way:
                units.append(_Unit(li, 1, o, da, D - da, 0, da))
                o += 2 * spec.n_lin
    for u in [None] + units:
        if len(lead if u is None else u.ops) > 4:
            raise NotImplementedError("more than 4 ActNorm / mixing layers between two coupling layers")
    return lead, units


def _fill_ops(arr: Any, ops: list[tuple[int, int]], params: Sequence[torch.Tensor], saves: list[Any],
              grads: list[Any] | None) -> int:
    for k, (typ, po) in enumerate(ops):
        arr[k].type = typ
        arr[k].p0 = params[po].data_ptr()
        if typ == _cabi.GLUE_ACTNORM:
            arr[k].p1 = params[po + 1].data_ptr()
            arr[k].save = saves[k].data_ptr()
            if grads is not None:
                arr[k].g0, arr[k].g1 = grads[po].data_ptr(), grads[po + 1].data_ptr()
    return len(ops)


def _call(name: str, args: Any, dev: torch.device) -> None:
    _cabi.check(getattr(_cabi.lib(), name)(C.byref(args), dev.index or 0, _stream(dev)), name)


def _pitch(n: int) -> int:
    return (n + 3) // 4 * 4


def _bwd_descs(u: Any, ws: Sequence[torch.Tensor], L: int, Cn: int, bwd: list[_Img]) -> list[tuple]:
    """img_pack descriptors of a network's weights in the data-gradient orientation (rows = input index)."""
    w1 = ws[0]
    return [(w1, u.din, 1, w1.stride(0), Cn, w1.shape[0], bwd[0])] + \
           [(ws[l], 0, 1, ws[l].stride(0), ws[l].shape[1], ws[l].shape[0], bwd[l]) for l in range(1, L)]


class _StackFn(torch.autograd.Function):
    """z, log|det J| = stack(y, h; parameters) with a hand-written backward.

    Per conditioner network the dependency chain is  pre (first Linear on the own half + P) -> hidden GEMMs ->
    post (last Linear, coupling, following ActNorm / mixing): L + 1 launches forward and backward.  Everything off
    that chain -- the condition projections P = h W1h^T + b1 of all networks, the weight / bias gradients and the
    gradient w.r.t. h -- runs on a second stream and only joins at the end.
    """

    @staticmethod
    def forward(ctx, spec: _Spec, y: torch.Tensor, h: torch.Tensor, *params: torch.Tensor):
        D, L = spec.size, spec.n_lin - 1
        y = y.contiguous().float()
        h = h.contiguous().float()
        B, Cn = y.shape[0], h.shape[1]
        dev = y.device
        lead, units = _plan(spec)
        main, side = torch.cuda.current_stream(dev), _side_stream(dev)
        ld = torch.zeros(B, device=dev)
        new = lambda *shape: torch.empty(*shape, device=dev)

        # off the chain, per network: operand images of its parameters (they changed in the last optimizer step), in
        # the forward (rows = out) and the data-gradient (rows = in) orientation, then the condition projection P
        side.wait_stream(main)
        n_units = len(units)
        ahead = _PACK_AHEAD if _PACK_AHEAD > 0 else n_units
        P: list[Any] = [None] * n_units
        p_ready: list[Any] = [None] * n_units
        wimg: list[Any] = [None] * n_units
        with torch.cuda.stream(side):
            # images this call saves for its backward: h (d W1h) and every hidden activation but the last of each network
            lease = _PoolLease(dev, [(B, Cn)] + [(B, params[u.w0 + l].shape[0]) for u in units for l in range(L - 1)])
            h_img, act_pool = lease.images[0], lease.images[1:]
            _pack_images([(h, 0, h.stride(0), 1, B, Cn, h_img)], dev)

        def prepare(idx: int, gate: Any) -> None:
            """Side stream: the forward-orientation images of network idx (all its images when nothing is deferred) and
            its condition projection, once the chain has reached the network `gate` was recorded at."""
            u = units[idx]
            with torch.cuda.stream(side):
                if gate is not None:
                    side.wait_event(gate)
                ws = params[u.w0: u.w0 + spec.n_lin]
                w1, b1 = ws[0], params[u.w0 + spec.n_lin]
                H1, pitch1 = w1.shape[0], w1.stride(0)
                fwd = [_img(dev, ("wf", w1.data_ptr()), H1, Cn, owner=w1)] + \
                      [_img(dev, ("wf", w.data_ptr()), *w.shape, owner=w) for w in ws[1:L]]
                bwd = [_img(dev, ("wb", w1.data_ptr()), Cn, H1, owner=w1)] + \
                      [_img(dev, ("wb", w.data_ptr()), w.shape[1], w.shape[0], owner=w) for w in ws[1:L]]
                descs = [(w1, u.din, pitch1, 1, H1, Cn, fwd[0])] + \
                        [(ws[l], 0, ws[l].stride(0), 1, ws[l].shape[0], ws[l].shape[1], fwd[l]) for l in range(1, L)]
                if _PACK_AHEAD <= 0:                   # the data-gradient orientation too (else: packed by the backward)
                    descs += _bwd_descs(u, ws, L, Cn, bwd)
                _pack_images(descs, dev)
                Pu = new(B, _pitch(H1))
                _gemm(None, None, None, None, Pu, B, H1, Cn, epi=_cabi.EPI_BIAS, bias=b1, split_k=1, a_img=h_img, b_img=fwd[0])
                ev = torch.cuda.Event()
                ev.record(side)
                P[idx], p_ready[idx], wimg[idx] = Pu, ev, (fwd, bwd)

        for idx in range(min(ahead, n_units)):
            prepare(idx, None)

        def glue_only(y_in, ops):
            a = _cabi.TrainPostArgs()
            y_out, saves = new(B, D), [new(B, D) if t == _cabi.GLUE_ACTNORM else None for t, _ in ops]
            a.B, a.D = B, D
            a.y_in, a.y_out, a.ld = y_in.data_ptr(), y_out.data_ptr(), ld.data_ptr()
            a.n_ops = _fill_ops(a.ops, ops, params, saves, None)
            _call("bcnf_train_post", a, dev)
            return y_out, saves

        lead_saves: list[Any] = []
        if lead:
            y, lead_saves = glue_only(y, lead)
        saved_units = []
        for ui, u in enumerate(units):
            if ui + ahead < n_units:
                gate = torch.cuda.Event()
                gate.record(main)
                prepare(ui + ahead, gate)
            ws = params[u.w0: u.w0 + spec.n_lin]
            bs = params[u.w0 + spec.n_lin: u.w0 + 2 * spec.n_lin]
            widths = [w.shape[0] for w in ws[:-1]]
            pre = [new(B, _pitch(n)) for n in widths]
            act = [new(B, _pitch(n)) for n in widths]
            main.wait_event(p_ready[ui])
            a = _cabi.TrainPreArgs()
            a.y, a.y_pitch, a.B, a.D, a.src0, a.din = y.data_ptr(), D, B, D, u.src0, u.din
            a.W1, a.w1_pitch, a.P, a.p_pitch, a.H = ws[0].data_ptr(), ws[0].stride(0), P[ui].data_ptr(), P[ui].stride(0), widths[0]
            a.pre, a.act, a.pitch = pre[0].data_ptr(), act[0].data_ptr(), pre[0].stride(0)
            a.seed, a.layer_uid, a.p_drop, a.seed_ptr = spec.seed, layer_uid(u.li, u.net, 0), spec.p_drop, spec.seed_ptr
            # images of the activations: the A operand of the next forward GEMM now, the B operand of the weight-gradient
            # GEMM in the backward pass.  They are part of what this call saves for its backward, so they belong to the
            # call (allocated zeroed: the padding must be zero), not to a shape-keyed cache
            aimg = act_pool[ui * (L - 1): (ui + 1) * (L - 1)] + [None]
            if L > 1:
                a.act_img, a.img_plane, a.img_rpad = aimg[0].ptr, aimg[0].plane, aimg[0].rpad
            _call("bcnf_train_pre", a, dev)
            for l in range(1, L):
                # act[l] = dropout(gelu(act[l-1] W_l^T + b_l)): nn.Linear, nn.GELU, nn.Dropout (cnf.py:79-83)
                _gemm(None, None, None, None, act[l], B, widths[l], widths[l - 1],
                      epi=_cabi.EPI_BIAS_GELU_DROP, bias=bs[l], save=pre[l], seed=spec.seed, uid=layer_uid(u.li, u.net, l),
                      p=spec.p_drop, seed_ptr=spec.seed_ptr, split_k=1, a_img=aimg[l - 1], b_img=wimg[ui][0][l],
                      c_img=aimg[l] if l < L - 1 else None)
            ls, ydst = new(B, u.dout), new(B, u.dout)
            y_out = new(B, D)
            op_saves = [new(B, D) if t == _cabi.GLUE_ACTNORM else None for t, _ in u.ops]
            a = _cabi.TrainPostArgs()
            a.a, a.a_pitch, a.Wout, a.bout = act[L - 1].data_ptr(), act[L - 1].stride(0), ws[L].data_ptr(), bs[L].data_ptr()
            a.B, a.D, a.H, a.dst0, a.dout = B, D, widths[L - 1], u.dst0, u.dout
            a.y_in, a.y_out, a.ld = y.data_ptr(), y_out.data_ptr(), ld.data_ptr()
            a.ls_save, a.ydst_save = ls.data_ptr(), ydst.data_ptr()
            a.n_ops = _fill_ops(a.ops, u.ops, params, op_saves, None)
            _call("bcnf_train_post", a, dev)
            saved_units.append((y, pre, act, ls, ydst, op_saves, aimg))
            y = y_out
        main.wait_stream(side)
        ctx.spec, ctx.params, ctx.h = spec, params, h
        # The parameters are held as plain attributes (their bf16 images, not the tensors, are what the backward reads),
        # so autograd's own version check does not see them: remember the versions and refuse a backward after an
        # in-place update (optimizer.step between forward and backward) instead of returning wrong gradients.
        ctx.param_versions = [t._version for t in params]
        ctx.plan = (lead, units, lead_saves, saved_units)
        ctx.wimg = wimg
        ctx.h_img = h_img
        ctx.lease = lease
        ctx.keep = P
        return y, ld

    @staticmethod
    def backward(ctx, dz: torch.Tensor, dld: torch.Tensor):
        spec, params, h = ctx.spec, ctx.params, ctx.h
        if not torch.cuda.is_current_stream_capturing():
            for i, (t, v) in enumerate(zip(params, ctx.param_versions)):
                if t._version != v:
                    raise RuntimeError(
                        f"one of the variables needed for gradient computation has been modified by an inplace operation: "
                        f"parameter {i} of the coupling stack is at version {t._version}; expected version {v} "
                        "(an optimizer step between forward and backward?)")
        lead, units, lead_saves, saved_units = ctx.plan
        D, L = spec.size, spec.n_lin - 1
        dz = dz.contiguous().float()
        B, Cn = dz.shape[0], h.shape[1]
        dev = dz.device
        dld = dld.contiguous().float() if dld is not None else torch.zeros(B, device=dev)
        main, sides = torch.cuda.current_stream(dev), _side_streams(dev)
        new = lambda *shape: torch.empty(*shape, device=dev)
        needs = ctx.needs_input_grad
        grads: list[Any] = [None] * len(params)
        # every gradient that is accumulated with atomics (ActNorm parameters, fused bias column sums) starts from
        # zero: one buffer, one fill
        n_zero = sum(params[po].numel() + params[po + 1].numel() for ops in [lead] + [u.ops for u in units]
                     for typ, po in ops if typ == _cabi.GLUE_ACTNORM)
        n_zero += sum(params[u.w0 + spec.n_lin + j].numel() for u in units for j in range(spec.n_lin))
        zero_buf, zero_at = torch.zeros(n_zero, device=dev), 0

        sink = _SINK if _SINK is not None and _SINK.active else None
        direct = [False] * len(params)          # gradient written straight into the parameter's .grad view: autograd gets None

        def zeros_like(t):
            nonlocal zero_at
            v = sink.view(t) if sink is not None else None      # (the sink's buffer was zeroed when the step began)
            if v is not None:
                return v
            v = zero_buf[zero_at: zero_at + t.numel()].view(t.shape)
            zero_at += t.numel()
            return v

        def empty_like(t):
            v = sink.view(t) if sink is not None else None
            return v if v is not None else torch.empty_like(t)

        if sink is not None:
            for i, t in enumerate(params):
                direct[i] = sink.view(t) is not None
            buckets = {u0: (lo, hi) for u0, lo, hi in sink.bucket_bounds(units, params)}

        for ops in [lead] + [u.ops for u in units]:
            for typ, po in ops:
                if typ == _cabi.GLUE_ACTNORM:
                    grads[po], grads[po + 1] = zeros_like(params[po]), zeros_like(params[po + 1])
        dh = new(B, Cn)
        grad_lease = _PoolLease(dev, [(B, params[u.w0 + l].shape[0]) for u in units for l in range(L)])
        grad_pool = grad_lease.images
        keep: list[Any] = [grad_lease]                                 # everything the side streams read stays alive until the join
        for sd in sides:
            sd.wait_stream(main)
        # data-gradient orientation of the parameter images, packed _PACK_AHEAD networks ahead of the chain on a stream
        # of their own (the side streams carry the weight-gradient GEMMs of earlier networks)
        n_units = len(units)
        bwd_ready: list[Any] = [None] * n_units
        pk = _pack_stream(dev) if _PACK_AHEAD > 0 else None

        def pack_bwd(idx: int, gate: Any) -> None:
            uu = units[idx]
            with torch.cuda.stream(pk):
                if gate is not None:
                    pk.wait_event(gate)
                _pack_images(_bwd_descs(uu, params[uu.w0: uu.w0 + spec.n_lin], L, Cn, ctx.wimg[idx][1]), dev)
                ev = torch.cuda.Event()
                ev.record(pk)
                bwd_ready[idx] = ev

        if pk is not None:
            pk.wait_stream(main)
            for idx in range(n_units - 1, max(n_units - 1 - _PACK_AHEAD, -1), -1):
                pack_bwd(idx, None)
        first_dh = True
        rr = 0
        for ui in range(len(units) - 1, -1, -1):
            u = units[ui]
            if pk is not None:
                if ui - _PACK_AHEAD >= 0:
                    gate = torch.cuda.Event()
                    gate.record(main)
                    pack_bwd(ui - _PACK_AHEAD, gate)
                main.wait_event(bwd_ready[ui])
            y_in, pre, act, ls, ydst, op_saves, aimg = saved_units[ui]
            ws = params[u.w0: u.w0 + spec.n_lin]
            widths = [w.shape[0] for w in ws[:-1]]
            no = 2 * u.dout
            d_pre = [new(B, _pitch(n)) for n in widths]
            d_o, dz_new = new(B, no), new(B, D)
            dws = [empty_like(w) for w in ws]
            dbs = [zeros_like(params[u.w0 + spec.n_lin + j]) for j in range(spec.n_lin)]
            a = _cabi.TrainPostBwdArgs()
            a.dz_in, a.dz_out, a.dld = dz.data_ptr(), dz_new.data_ptr(), dld.data_ptr()
            a.B, a.D, a.H, a.dst0, a.dout = B, D, widths[L - 1], u.dst0, u.dout
            a.ls_save, a.ydst_save, a.Wout = ls.data_ptr(), ydst.data_ptr(), ws[L].data_ptr()
            a.pre, a.pitch, a.d_o, a.d_pre = pre[L - 1].data_ptr(), pre[L - 1].stride(0), d_o.data_ptr(), d_pre[L - 1].data_ptr()
            a.seed, a.layer_uid, a.p_drop, a.seed_ptr = spec.seed, layer_uid(u.li, u.net, L - 1), spec.p_drop, spec.seed_ptr
            a.n_ops = _fill_ops(a.ops, u.ops, params, op_saves, grads)
            # images of d pre[l]: A operand of the next data-gradient GEMM on this stream and of the weight-gradient
            # GEMMs on the side streams (one per network and layer)
            gimg = grad_pool[ui * L: (ui + 1) * L]
            a.dpre_img, a.img_plane, a.img_rpad = gimg[L - 1].ptr, gimg[L - 1].plane, gimg[L - 1].rpad
            _call("bcnf_train_post_bwd", a, dev)
            for l in range(L - 1, 0, -1):
                # d pre[l-1] = (d pre[l] W_l) * gelu'(pre[l-1]) * mask[l-1]; its column sums are the bias gradient of Linear l-1
                _gemm(None, None, None, None, d_pre[l - 1], B, widths[l - 1], widths[l],
                      epi=_cabi.EPI_DGELU_DROP, saved=pre[l - 1], seed=spec.seed, uid=layer_uid(u.li, u.net, l - 1),
                      p=spec.p_drop, seed_ptr=spec.seed_ptr, colsum=dbs[l - 1], split_k=1, a_img=gimg[l],
                      b_img=ctx.wimg[ui][1][l], c_img=gimg[l - 1])
            a = _cabi.TrainPreBwdArgs()
            a.d_pre, a.pitch, a.W1, a.w1_pitch = d_pre[0].data_ptr(), d_pre[0].stride(0), ws[0].data_ptr(), ws[0].stride(0)
            a.B, a.D, a.H, a.src0, a.din, a.dz = B, D, widths[0], u.src0, u.din, dz_new.data_ptr()
            _call("bcnf_train_pre_bwd", a, dev)
            done = torch.cuda.Event()
            done.record(main)
            # ---- off the chain: parameter gradients and d h, spread over the side streams ----
            for sd in sides:
                sd.wait_event(done)
            for l in range(L - 1, 0, -1):
                with torch.cuda.stream(sides[rr % len(sides)]):
                    # dW_l = d pre[l]^T act[l-1]: both images read MN-major (contraction over the batch rows)
                    _gemm(None, None, None, None, dws[l], widths[l], widths[l - 1], B, split_k=1, a_img=gimg[l],
                          b_img=aimg[l - 1], mn=True)
                rr += 1
            w1g = dws[0]
            with torch.cuda.stream(sides[rr % len(sides)]):
                # first Linear: columns [din, din + C) against h
                _gemm(None, None, None, None, w1g[:, u.din:], widths[0], Cn, B, split_k=1, c_stride=w1g.stride(0),
                      a_img=gimg[0], b_img=ctx.h_img, mn=True)
            rr += 1
            with torch.cuda.stream(sides[0]):
                _colsum(d_pre[L - 1], dbs[L - 1], cols=widths[L - 1])
                _colsum(d_o, dbs[L])
                # dW_L = d_o^T act[L-1]
                _gemm(d_o, (1, no), act[L - 1], (act[L - 1].stride(0), 1), dws[L], no, widths[L - 1], B, split_k=1)
                # first Linear: columns [0, din) against the own half of y
                ys = y_in[:, u.src0:]
                _gemm(d_pre[0], (1, d_pre[0].stride(0)), ys, (y_in.stride(0), 1), w1g, widths[0], u.din, B, split_k=1,
                      c_stride=w1g.stride(0))
                if needs[2]:
                    _gemm(None, None, None, None, dh, B, Cn, widths[0], beta=0.0 if first_dh else 1.0, split_k=1,
                          a_img=gimg[0], b_img=ctx.wimg[ui][1][0])
                    first_dh = False
            for j in range(spec.n_lin):
                grads[u.w0 + j] = dws[j]
                grads[u.w0 + spec.n_lin + j] = dbs[j]
            keep += [d_pre, d_o, dz]
            dz = dz_new
            if sink is not None and ui in buckets and ui != 0:
                # every gradient of units >= ui is enqueued: all-reduce their range underneath the rest of the backward
                sink.reduce_range(*buckets[ui], [main] + list(sides))
        if lead:
            dz_new = new(B, D)
            a = _cabi.TrainPostBwdArgs()
            a.dz_in, a.dz_out, a.dld, a.B, a.D = dz.data_ptr(), dz_new.data_ptr(), dld.data_ptr(), B, D
            a.n_ops = _fill_ops(a.ops, lead, params, lead_saves, grads)
            _call("bcnf_train_post_bwd", a, dev)
            keep.append(dz)
            dz = dz_new
        for sd in sides:
            main.wait_stream(sd)
        if pk is not None:
            main.wait_stream(pk)
        if sink is not None:
            sink.reduce_range(*buckets[0], [main])          # the first bucket (incl. the leading ActNorm) closes the backward
        del keep
        out_params = [g if needs[3 + i] and not direct[i] else None for i, g in enumerate(grads)]
        return (None, dz if needs[1] else None, dh if needs[2] and not first_dh else None, *out_params)



# ------------------------------------------------------------------------------------------
# gradient sink: parameter gradients as views of ONE flat buffer, all-reduced bucket by bucket during the backward
# ------------------------------------------------------------------------------------------
class _GradSink:
    """Data-parallel gradient plumbing of the stack (48.8 M fp32 = 195 MB per step).

    Every stack parameter's ``.grad`` is a view of one flat buffer.  While the sink is active the stack's backward
    writes each gradient straight into its view (and hands autograd ``None`` for it): no per-parameter AccumulateGrad
    nodes, no gather into / scatter out of a communication buffer.  The flat buffer is cut into ``n_buckets``
    contiguous ranges of coupling layers; the backward walks the layers from last to first, so a range is complete
    while earlier layers are still back-propagating, and its NCCL all-reduce runs on a side stream underneath them
    (captured with the rest of the step when the Trainer replays a CUDA graph).  The loss is pre-scaled by 1 / world,
    so the SUM all-reduce yields the mean without another pass over the buffer.
    """

    def __init__(self, params: list[torch.Tensor], process_group: Any, n_buckets: int = 4,
                 adopt: "FlatAdam | None" = None) -> None:
        self.group = process_group
        self.n_buckets = n_buckets
        self.params = [t for t in params if t.requires_grad]
        dev = self.params[0].device
        if adopt is not None:
            # the optimizer's flat gradient blob IS the sink: its first `stack_end` elements hold the stack's gradients
            # in this order and with this alignment
            if [id(t) for t in adopt.params[: len(self.params)]] != [id(t) for t in self.params]:
                raise ValueError("FlatAdam was built for another model (its parameter order does not start with the stack's)")
            self.offs, self.total, self.flat = adopt.offs, adopt.stack_end, adopt.flat_g
        else:
            offs, at = {}, 0
            for t in self.params:
                offs[id(t)] = at
                at += _align64(t.numel())                   # 256-byte aligned views (vector stores of the GEMM epilogues)
            self.offs, self.total = offs, at
            self.flat = torch.zeros(at, device=dev)
            for t in self.params:
                t.grad = self.view(t)
        self.comm = torch.cuda.Stream(device=dev)
        self.active = False
        self._pending = False
        # With an adopted FlatAdam the optimizer step of a bucket follows its all-reduce on the communication stream, while
        # the backward of the earlier coupling layers (which read none of those parameters any more) is still running:
        # the 0.5 ms pass over 4 x 195 MB leaves the end of the step.  Armed by the Trainer for steps it drives as a whole
        # (train_batch); a bare _backward() -- gradient accumulation, tests -- leaves the parameters alone.
        self.adam = adopt if _EARLY_ADAM else None
        self.armed = False

    def view(self, t: torch.Tensor) -> torch.Tensor | None:
        o = self.offs.get(id(t))
        return None if o is None else self.flat[o: o + t.numel()].view(t.shape)

    def begin(self) -> None:
        self.flat[: self.total].zero_()                     # bias / ActNorm gradients are accumulated with atomics
        self.active, self._pending = True, False
        if self.adam is not None and self.armed:
            self.adam.begin_early_step()

    def bucket_bounds(self, units: list[Any], params: list[torch.Tensor]) -> list[tuple[int, int, int]]:
        """(first unit, element lo, element hi) of each bucket; a bucket is complete when its first unit is done."""
        n = len(units)
        nb = max(1, min(self.n_buckets, n))
        firsts = sorted({(b * n) // nb for b in range(nb)})
        starts = [0 if u == 0 else self.offs[id(params[units[u].w0])] for u in firsts]
        ends = starts[1:] + [self.total]
        return [(u, lo, hi) for u, lo, hi in zip(firsts, starts, ends)]

    def reduce_range(self, lo: int, hi: int, streams: list[torch.cuda.Stream]) -> None:
        """All-reduce flat[lo:hi] on the communication stream once everything enqueued so far on `streams` is done."""
        early = self.adam is not None and self.armed
        if (self.group is None and not early) or hi <= lo:
            return
        for st in streams:
            self.comm.wait_stream(st)
        with torch.cuda.stream(self.comm):
            if self.group is not None:
                import torch.distributed as dist
                dist.all_reduce(self.flat[lo:hi], group=self.group)
            if early:
                self.adam.step_range(lo, hi)
        self._pending = True

    def end(self, main: torch.cuda.Stream) -> None:
        if self._pending:
            main.wait_stream(self.comm)
        self.active, self._pending = False, False


_SINK: _GradSink | None = None      # set by Trainer for the duration of its backward


def _align64(n: int) -> int:
    return (n + 63) // 64 * 64


class FlatAdam(torch.optim.Optimizer):
    """``torch.optim.Adam`` (no amsgrad) over ONE flat blob: SURVEY.md section 8f-3, reference step trainer.py:271.

    Parameters, gradients and both moments of the model live in four flat fp32 buffers of one layout (every tensor
    a 256-byte-aligned view; the coupling stack's parameters first, in the order the stack's backward produces their
    gradients, so that the Trainer's gradient sink and its all-reduce buckets are ranges of the same buffer).  A step
    is one launch (``bcnf_adam_flat``: one pass over 4 x n floats) instead of a multi-tensor launch per chunk of
    tensors; learning rate, betas, eps, weight decay and the step count live on the device, so a captured step replays
    with the current values (``sync_hyper`` pushes a changed ``param_groups[0]["lr"]``).

    A ``torch.optim.Optimizer`` (one parameter group; learning-rate schedulers work on it) whose ``state`` is the
    three flat tensors.  CUDA models without cuDNN RNN modules (their weights must stay in cuDNN's own flat buffer).
    """

    def __init__(self, model: torch.nn.Module, lr: float = 1e-3, betas: tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 0.0) -> None:
        if any(isinstance(m, torch.nn.RNNBase) for m in model.modules()):
            raise NotImplementedError("FlatAdam: models with cuDNN RNN modules are not supported (use torch.optim.Adam)")
        stack = [t for t in stack_parameters(model) if t.requires_grad]
        seen = {id(t) for t in stack}
        self.params = stack + [t for t in model.parameters() if t.requires_grad and id(t) not in seen]
        if not self.params or not all(t.is_cuda and t.dtype == torch.float32 for t in self.params):
            raise ValueError("FlatAdam needs fp32 CUDA parameters")
        dev = self.params[0].device
        self.offs, at = {}, 0
        for i, t in enumerate(self.params):
            if i == len(stack):
                self.stack_end = at
            self.offs[id(t)] = at
            at += _align64(t.numel())
        if len(stack) == len(self.params):
            self.stack_end = at
        self.total = at
        self.flat_p = torch.zeros(at, device=dev)
        self.flat_g = torch.zeros(at, device=dev)
        self.exp_avg = torch.zeros(at, device=dev)
        self.exp_avg_sq = torch.zeros(at, device=dev)
        self.step_t = torch.zeros((), device=dev)
        with torch.no_grad():
            for t in self.params:
                o, n = self.offs[id(t)], t.numel()
                view = self.flat_p[o: o + n].view(t.shape)
                view.copy_(t)
                t.data = view
                t.grad = self.flat_g[o: o + n].view(t.shape)
        super().__init__(self.params, {"lr": lr, "betas": tuple(betas), "eps": eps, "weight_decay": weight_decay,
                                       "capturable": True})
        self.state["flat"] = {"step": self.step_t, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq}
        self._hyper_host: tuple | None = None
        self.hyper = torch.zeros(5, device=dev)
        self.sync_hyper()

    def sync_hyper(self) -> None:
        g = self.param_groups[0]
        host = (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]))
        if host != self._hyper_host:
            self.hyper.copy_(torch.tensor(host, dtype=torch.float32), non_blocking=True)
            self._hyper_host = host

    def zero_grad(self, set_to_none: bool = True) -> None:      # (the views stay: "none" would detach them from the blob)
        self.flat_g.zero_()

    _early = False      # this step's count is advanced and the stack's range is being applied bucket by bucket (_GradSink)

    @torch.no_grad()
    def begin_early_step(self) -> None:
        """Called by the Trainer's gradient sink at the start of a backward pass it will follow with step(): advances the
        step count; the ranges of the stack's parameters are then applied by step_range as their gradients become final
        and step() only covers what is left (the feature network's parameters)."""
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        self.step_t.add_(1.0)
        self._early = True

    @torch.no_grad()
    def step_range(self, lo: int, hi: int) -> None:
        """Adam on elements [lo, hi) of the blob, on the current stream (the step count is not touched)."""
        if hi <= lo:
            return
        dev = self.flat_p.device
        rc = _cabi.lib().bcnf_adam_flat(self.flat_p[lo:].data_ptr(), self.flat_g[lo:].data_ptr(), self.exp_avg[lo:].data_ptr(),
                                        self.exp_avg_sq[lo:].data_ptr(), hi - lo, self.hyper.data_ptr(),
                                        self.step_t.data_ptr(), dev.index or 0, _stream(dev))
        _cabi.check(rc, "bcnf_adam_flat")

    @torch.no_grad()
    def step(self, closure: Any = None) -> None:
        if self._early:                                     # the stack's range went out during the backward
            self._early = False
            self.step_range(self.stack_end, self.total)
            return
        if not torch.cuda.is_current_stream_capturing():
            self.sync_hyper()
        self.step_t.add_(1.0)
        self.step_range(0, self.total)

    def state_dict(self) -> dict:
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        return {"flat": {k: v.detach().clone() for k, v in self.state["flat"].items()}, "param_group": g,
                "layout": [(tuple(t.shape), self.offs[id(t)]) for t in self.params]}

    def load_state_dict(self, sd: dict) -> None:
        if [(tuple(t.shape), self.offs[id(t)]) for t in self.params] != [(tuple(sh), o) for sh, o in sd["layout"]]:
            raise ValueError("FlatAdam.load_state_dict: parameter layout differs")
        with torch.no_grad():
            for k, v in sd["flat"].items():
                self.state["flat"][k].copy_(v)
        self.param_groups[0].update(sd["param_group"])
        self.sync_hyper()


class _NllFn(torch.autograd.Function):
    """inn_nll_loss(z, logdet) (utils.py:49-53) and the gradients it sends back, one launch (``bcnf_train_nll``)."""

    @staticmethod
    def forward(ctx, z: torch.Tensor, ld: torch.Tensor):
        z, ld = z.contiguous(), ld.contiguous()
        dev = z.device
        loss = torch.empty((), device=dev)
        dz, dld = torch.empty_like(z), torch.empty_like(ld)
        rc = _cabi.lib().bcnf_train_nll(z.data_ptr(), ld.data_ptr(), z.shape[0], z.shape[1], loss.data_ptr(),
                                        dz.data_ptr(), dld.data_ptr(), dev.index or 0, _stream(dev))
        _cabi.check(rc, "bcnf_train_nll")
        ctx.save_for_backward(dz, dld)
        return loss

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        dz, dld = ctx.saved_tensors
        return dz * g, dld * g


def fused_nll(z: torch.Tensor, log_det_J: torch.Tensor) -> torch.Tensor:
    """``inn_nll_loss(z, log_det_J)`` (mean reduction) as one kernel forward; fp32 CUDA (B, D) / (B,) tensors."""
    return _NllFn.apply(z, log_det_J)

def stack_parameters(model: Any) -> list[torch.Tensor]:
    """The parameters of ``model.layers`` in the order stack_forward_train hands them to the autograd function."""
    from .cnf import ActNorm, ConditionalAffineCouplingLayer, OrthonormalTransformation
    out: list[torch.Tensor] = []
    for layer in getattr(model, "layers", []):
        if isinstance(layer, ActNorm):
            out += [layer.scale, layer.bias]
        elif isinstance(layer, OrthonormalTransformation):
            out.append(layer.orthonormal_matrix)
        elif isinstance(layer, ConditionalAffineCouplingLayer):
            for net in ([layer.nn_a, layer.nn_b] if layer.two_way else [layer.nn_a]):
                lin = net.linears()
                out += [m.weight for m in lin] + [m.bias for m in lin]
    return out


def stack_forward_train(model: Any, y: torch.Tensor, h: torch.Tensor, seed: int | None = None,
                        seed_word: torch.Tensor | None = None):
    """Training-mode forward of ``model.layers``: returns (z, log|det J|), both with autograd history.

    ``seed_word``: optional device int64 tensor whose value is XORed into the dropout seed inside the kernels;
    a CUDA-graph-captured step bumps it on the device so that every replay draws fresh masks.
    """
    from .cnf import ActNorm, ConditionalAffineCouplingLayer, OrthonormalTransformation
    kinds, params = [], []
    n_lin = len(model.nested_sizes) + 1
    for layer in model.layers:
        if isinstance(layer, ActNorm):
            kinds.append("actnorm")
            params += [layer.scale, layer.bias]
        elif isinstance(layer, OrthonormalTransformation):
            kinds.append("ortho")
            params.append(layer.orthonormal_matrix)
        elif isinstance(layer, ConditionalAffineCouplingLayer):
            kinds.append("coupling")
            for net in ([layer.nn_a, layer.nn_b] if layer.two_way else [layer.nn_a]):
                lin = net.linears()
                params += [m.weight for m in lin] + [m.bias for m in lin]
        else:
            raise ValueError(f"Layer must be an instance of ConditionalInvertibleLayer or InvertibleLayer, but got {type(layer)}")
    if seed_word is None:
        seed_word = getattr(model, "_dropout_seed_word", None)
    if seed is None:
        # eager mode consumes the CPU generator, like nn.Dropout would; under graph capture the base seed is fixed
        seed = 0x5EED if seed_word is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
    p = float(model.dropout) if model.training else 0.0
    spec = _Spec(kinds, n_lin, model.two_way, model.size, model.n_conditions, p, seed, seed_word)
    dev = torch.device(model.device)
    if dev.type != "cuda":
        raise RuntimeError(f"bcnf_b200 runs on CUDA devices only (got {dev}); there is no CPU path. "
                           "Move the model with .to('cuda').")
    for t in params:
        if t.device.type != "cuda" or t.dtype != torch.float32:
            raise RuntimeError("training needs float32 parameters on the model's CUDA device "
                               f"(got {t.dtype} on {t.device})")
    # torch.linalg.qr hands back a column-major Q: make every operand row-major (differentiable no-op otherwise)
    params = [t if t.is_contiguous() else t.contiguous() for t in params]
    return _StackFn.apply(spec, y.to(dev), h.to(dev), *params)


class Trainer:
    """The optimisation step of the reference Trainer (trainer.py:244-303) without its data / wandb plumbing.

    ``train_batch`` = zero_grad -> forward(log_det_J, return_features) -> [hybrid MSE head] -> NLL ->
    backward -> optimizer.step -> clip_grad_norm_ (after the step, as the reference does: it never
    affects an update, trainer.py:273-275).  Returns Python floats like the reference (3 syncs).

    ``cuda_graph=True`` captures zero_grad + forward + backward + optimizer.step of one batch shape into a CUDA
    graph and replays it (the step is launch-bound at the reference's batch size of 256); it needs a
    capturable optimizer (``torch.optim.Adam(..., capturable=True)``).

    Data parallel training either wraps the model in ``DistributedDataParallel`` (eager
    steps only), or -- ``process_group=`` given, model NOT wrapped -- keeps the whole step in the graph: the
    stack's gradients are views of one flat buffer that the backward writes directly (``_GradSink``), all-reduced
    bucket by bucket on a communication stream while the backward of the earlier blocks runs.  Parameters are
    broadcast from rank 0 when the trainer is built, as DistributedDataParallel does.  With ``FlatAdam`` as the
    optimizer that flat buffer is the optimizer's gradient blob and the step is one launch.
    """

    def __init__(self, model: Any, optimizer: torch.optim.Optimizer, hybrid_weight: float = 0.0,
                 cuda_graph: bool = False, process_group: Any = None) -> None:
        from .utils import inn_nll_loss
        self.model, self.optimizer, self.hybrid_weight = model, optimizer, hybrid_weight
        self.loss_function = inn_nll_loss
        self.mse_loss = torch.nn.MSELoss()
        self.cuda_graph = cuda_graph
        self.process_group = process_group
        self._graphs: dict[tuple, dict[str, Any]] = {}      # one captured step per batch shape
        self._sink: _GradSink | None = None
        self._other: list[torch.Tensor] = []
        self._world = 1
        self._flat_opt = optimizer if isinstance(optimizer, FlatAdam) else None
        if self._flat_opt is not None:
            if hasattr(model, "module"):
                raise ValueError("FlatAdam owns the gradient buffers: pass the bare model (and process_group= for data parallel)")
            sp = [t for t in stack_parameters(model) if t.requires_grad]
            if sp and process_group is None:            # single GPU: the stack still writes its gradients into the blob
                self._sink = _GradSink(sp, None, adopt=self._flat_opt)
                sunk = {id(t) for t in self._sink.params}
                self._other = [t for t in model.parameters() if t.requires_grad and id(t) not in sunk]
        if cuda_graph:
            if hasattr(model, "module"):
                raise NotImplementedError("cuda_graph=True with a DistributedDataParallel wrapper is not supported: pass the "
                                          "bare model and process_group= instead")
            if not all(g.get("capturable", False) for g in optimizer.param_groups):
                raise ValueError("cuda_graph=True needs a capturable optimizer, e.g. torch.optim.Adam(..., capturable=True)")
        if process_group is not None:
            import torch.distributed as dist
            if hasattr(model, "module"):
                raise ValueError("process_group= replaces the DistributedDataParallel wrapper: pass the bare model")
            self._world = dist.get_world_size(process_group)
            # the stack's gradients live in one flat buffer and are all-reduced bucket by bucket during the backward
            # (CUDA models with a coupling stack; anything else -- the gloo tests of this plumbing -- takes the generic
            # one-buffer path of _allreduce_grads after the backward)
            sp = [t for t in stack_parameters(model) if t.requires_grad and t.is_cuda]
            if sp:
                self._sink = _GradSink(sp, process_group, adopt=self._flat_opt)
                sunk = {id(t) for t in self._sink.params}
                self._other = [t for t in model.parameters() if t.requires_grad and id(t) not in sunk]
            with torch.no_grad():
                src = dist.get_global_rank(process_group, 0)
                for t in list(model.parameters()) + list(model.buffers()):
                    if t.is_contiguous():
                        dist.broadcast(t, src=src, group=process_group)
                    else:                                 # torch.linalg.qr hands back a column-major Q
                        tmp = t.contiguous()
                        dist.broadcast(tmp, src=src, group=process_group)
                        t.copy_(tmp)

    def close(self) -> None:
        """Drop the captured graph.  With process_group= the graph holds NCCL kernels of the group's communicator:
        call this (and synchronize) before ``torch.distributed.destroy_process_group()``, which otherwise waits for it."""
        self._graphs.clear()
        self._flat = None
        net = self._net()
        if getattr(net, "_dropout_seed_word", None) is not None:
            net._dropout_seed_word = None       # eager train-mode forwards draw their own seeds again
        clear_caches()

    def _allreduce_grads(self) -> None:
        """Average the gradients over the process group through one flat buffer (one NCCL call)."""
        import torch.distributed as dist
        grads = [p.grad for p in self._net().parameters() if p.grad is not None]
        if not grads or self._world == 1:
            return
        flat = getattr(self, "_flat", None)
        n = sum(g.numel() for g in grads)
        if flat is None or flat.numel() != n:
            flat = self._flat = torch.empty(n, device=grads[0].device)
        views, at = [], 0
        for g in grads:
            views.append(flat[at: at + g.numel()].view(g.shape))
            at += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(flat, group=self.process_group)
        flat.mul_(1.0 / self._world)
        torch._foreach_copy_(grads, views)

    def _net(self) -> Any:
        return self.model.module if hasattr(self.model, "module") else self.model

    def _zero_grad(self) -> None:
        if self._sink is None:
            self.optimizer.zero_grad(set_to_none=True)
            return
        if self._flat_opt is not None:           # the other parameters' gradients are views of the same blob: one fill
            self._flat_opt.flat_g[self._flat_opt.stack_end:].zero_()
            return
        for t in self._other:                    # the stack's .grad views stay: the sink zeroes their buffer in one fill
            t.grad = None

    def _backward(self, loss: torch.Tensor) -> None:
        """loss.backward(), with the stack's gradients written into the sink and all-reduced underneath the backward."""
        global _SINK
        oc = getattr(self, "_oc", None)
        if self._sink is None:
            loss.backward()
            if oc is not None:
                oc.join(torch.cuda.current_stream(loss.device))
            if self.process_group is not None:
                self._allreduce_grads()
            return
        dev = self._sink.flat.device
        self._sink.begin()
        _SINK = self._sink
        try:
            (loss / self._world).backward()       # SUM all-reduce of pre-scaled gradients = their mean
        finally:
            _SINK = None
        self._sink.end(torch.cuda.current_stream(dev))
        if oc is not None:                       # the encoder's deferred weight / bias gradients (feature_network.OffChain)
            oc.join(torch.cuda.current_stream(dev))
        if self._world > 1:
            self._allreduce_other()

    def _allreduce_other(self) -> None:
        import torch.distributed as dist
        if self._flat_opt is not None:           # contiguous in the blob: no gather / scatter
            tail = self._flat_opt.flat_g[self._flat_opt.stack_end:]
            if tail.numel():
                dist.all_reduce(tail, group=self.process_group)
            return
        grads = [p.grad for p in self._other if p.grad is not None]
        if not grads:
            return
        n = sum(g.numel() for g in grads)
        flat = getattr(self, "_flat", None)
        if flat is None or flat.numel() != n:
            flat = self._flat = torch.empty(n, device=grads[0].device)
        views, at = [], 0
        for g in grads:
            views.append(flat[at: at + g.numel()].view(g.shape))
            at += g.numel()
        torch._foreach_copy_(views, grads)
        dist.all_reduce(flat, group=self.process_group)      # (already scaled by 1 / world through the loss)
        torch._foreach_copy_(grads, views)

    def _off_chain(self, dev: torch.device) -> Any:
        """Side streams for the encoder's deferred parameter gradients (feature_network.OffChain), or None."""
        # (a torch DistributedDataParallel wrapper needs every gradient to pass its autograd hooks: plain path there)
        if not _ENC_OFF_CHAIN or torch.device(dev).type != "cuda" or not self.model.training or hasattr(self.model, "module"):
            return None
        oc = getattr(self, "_oc", None)
        if oc is None:
            from .feature_network import OffChain
            oc = self._oc = OffChain(_side_streams(torch.device(dev)))
        return oc

    def _losses(self, y: torch.Tensor, *conditions: torch.Tensor):
        from . import feature_network as _fn
        net = self._net()
        dev = net.device
        _fn._OFF_CHAIN = self._off_chain(dev)
        try:
            z, h = self.model(y.to(dev), *[c.to(dev) for c in conditions], log_det_J=True, return_features=True)
        finally:
            _fn._OFF_CHAIN = None
        if self.hybrid_weight > 0:
            mse = self.mse_loss(net.prediction_head(h), y.to(dev))
        else:
            mse = torch.zeros((), device=z.device)
        if z.is_cuda and z.dtype == torch.float32 and z.ndim == 2:
            nll = fused_nll(z, net.log_det_J)             # the same value and gradients in one launch each way
        else:
            nll = self.loss_function(z, net.log_det_J)
        loss = (nll + mse * self.hybrid_weight) / (1 + self.hybrid_weight)
        return loss, nll, mse, z

    def _seed_word(self, net: Any, dev: torch.device) -> torch.Tensor:
        """Device-resident dropout counter, created ONCE per model: it keeps counting across graph captures (a reset
        would replay the same mask sequence after every new batch shape) and starts from a random base mixed with the
        rank, so that replicas draw different masks (nn.Dropout under DistributedDataParallel does too)."""
        w = getattr(net, "_dropout_seed_word", None)
        if w is None or w.device != dev:
            rank = 0
            if self.process_group is not None:
                import torch.distributed as dist
                rank = dist.get_rank(self.process_group)
            base = int(torch.randint(0, 2 ** 40, (1,)).item())            # consumes the CPU generator, like nn.Dropout
            w = net._dropout_seed_word = torch.tensor([base + (rank << 44)], dtype=torch.int64, device=dev)
        return w

    def _graphed_step(self, y: torch.Tensor, *conditions: torch.Tensor):
        net = self._net()
        dev = torch.device(net.device)
        shapes = (tuple(y.shape),) + tuple(tuple(c.shape) for c in conditions)
        st = self._graphs.get(shapes)
        if st is None:
            seed_word = self._seed_word(net, dev)
            st = {"shapes": shapes, "y": torch.empty_like(y, device=dev),
                  "c": [torch.empty_like(c, device=dev) for c in conditions]}
            st["y"].copy_(y)
            for d, c in zip(st["c"], conditions):
                d.copy_(c)
            # Warm-up off the capture stream (allocator, lazy initialisations incl. the optimizer's state tensors).
            # It must not count as training: parameters, optimizer state and the seed word are put back afterwards --
            # a partial last batch would otherwise give the model four extra Adam updates per epoch.
            params = [p for p in net.parameters()]
            snap_p = [p.detach().clone() for p in params]
            snap_o = {id(t): t.detach().clone() for stt in self.optimizer.state.values() for t in stt.values()
                      if torch.is_tensor(t)}
            snap_w = seed_word.clone()
            # Warm-up and capture run on ONE stream: autograd's AccumulateGrad nodes remember the stream they were created
            # on, they outlive the warm-up (model.log_det_J keeps its graph alive), and a captured backward that has to
            # synchronise with a different, uncaptured stream fails with cudaErrorStreamCaptureIsolation.
            side = getattr(self, "_cap_stream", None)
            if side is None:
                side = self._cap_stream = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                def warm_step():
                    # (a function of its own: nothing of the warm-up's autograd graph -- loss, nll, mse, z -- may outlive
                    # it.  The parameters' AccumulateGrad nodes remember the stream they were created on, and nodes
                    # kept alive by such a reference would make the captured backward synchronise with this
                    # (uncaptured) warm-up stream: cudaErrorStreamCaptureIsolation when no defined gradient reaches
                    # them, which is the case for every stack parameter behind the sink.)
                    self._zero_grad()
                    loss = self._losses(st["y"], *st["c"])[0]
                    self._backward_and_step(loss)
                    seed_word.add_(1)
                for _ in range(3):
                    warm_step()
                import gc
                gc.collect()
                with torch.no_grad():
                    for p, q in zip(params, snap_p):
                        p.copy_(q)
                    for stt in self.optimizer.state.values():
                        for t in stt.values():
                            if torch.is_tensor(t):
                                t.copy_(snap_o[id(t)]) if id(t) in snap_o else t.zero_()   # created by the warm-up: as new
                    seed_word.copy_(snap_w)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            self._zero_grad()
            # NCCL's watchdog thread polls events while the capture is open: keep the capture check thread-local
            mode = {"capture_error_mode": "thread_local"} if self.process_group is not None else {}
            with torch.cuda.graph(graph, stream=side, **mode):
                loss, nll, mse, _ = self._losses(st["y"], *st["c"])
                self._backward_and_step(loss)
                seed_word.add_(1)                       # fresh dropout masks on every replay
            st.update(loss=loss, nll=nll, mse=mse, graph=graph)
            self._graphs[shapes] = st
        st["y"].copy_(y, non_blocking=True)
        for d, c in zip(st["c"], conditions):
            d.copy_(c, non_blocking=True)
        if self._flat_opt is not None:
            self._flat_opt.sync_hyper()                 # a scheduler may have changed the learning rate since the capture
        st["graph"].replay()
        return st["loss"], st["nll"], st["mse"]

    def _backward_and_step(self, loss: torch.Tensor) -> None:
        """Backward pass + optimizer step of a step the Trainer drives as a whole: with FlatAdam the stack's parameters are
        updated bucket by bucket behind their gradient all-reduce, underneath the rest of the backward (_GradSink)."""
        if self._sink is not None:
            self._sink.armed = True
        try:
            self._backward(loss)
        except BaseException:
            if self._flat_opt is not None:
                self._flat_opt._early = False
            raise
        finally:
            if self._sink is not None:
                self._sink.armed = False
        self.optimizer.step()

    def train_batch_async(self, y: torch.Tensor, *conditions: torch.Tensor) -> torch.Tensor:
        """One training step without a host synchronisation: returns the device tensor [loss, nll, mse].  Steps enqueued
        back to back overlap the host-side launch of a step (a CUDA graph of ~1000 nodes: ~0.6 ms) with the previous one."""
        if self.cuda_graph:
            return torch.stack([t.detach().reshape(()) for t in self._graphed_step(y, *conditions)])
        self._zero_grad() if self._sink is not None else self.optimizer.zero_grad()
        loss, nll, mse, _ = self._losses(y, *conditions)
        self._backward_and_step(loss)
        torch.nn.utils.clip_grad_norm_(self._net().parameters(), max_norm=1.0)
        return torch.stack([t.detach().reshape(()) for t in (loss, nll, mse)])

    def train_batch(self, y: torch.Tensor, *conditions: torch.Tensor) -> tuple[float, float, float]:
        """trainer.py:244-277: one step, losses read back as Python floats (one device -> host copy)."""
        loss, nll, mse = self.train_batch_async(y, *conditions).tolist()
        return loss, nll, mse

    def validate_batch(self, y: torch.Tensor, *conditions: torch.Tensor):
        with torch.no_grad():
            loss, nll, mse, z = self._losses(y, *conditions)
        return loss.item(), nll.item(), mse.item(), z.mean(dim=0), z.std(dim=0)
