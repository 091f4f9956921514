"""Training step of the coupling stack (reference ``Trainer._train_batch``, src/bcnf/train/trainer.py:244-277).

``CondRealNVP_v2.forward`` in training mode routes here: one ``torch.autograd.Function`` for the
whole stack.  The conditioner's Linear -> GELU -> Dropout chains -- all of the FLOPs -- run forward and
backward on the fused SGEMM of ``bcnf_b200/csrc/train_ops.cuh`` through the C ABI
(``bcnf_train_gemm``): bias + GELU + dropout in the forward epilogue, gelu' * mask in the data-gradient
epilogue, weight gradients accumulated straight into tensors shaped like the parameters.  Dropout
masks are a counter-based hash of (seed, layer, row, column), regenerated in backward, never stored.
The D-wide glue between the GEMMs (tanh / exp / affine update, ActNorm, the D x D mixing) is a handful
of tiny torch ops on (B, D) tensors.

Because gradients are returned per parameter tensor, ``torch.optim`` and
``torch.nn.parallel.DistributedDataParallel`` (NCCL all-reduce of the gradients) work
unchanged on top.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Sequence

import torch

from . import _cabi

__all__ = ["stack_forward_train", "dropout_mask", "Trainer"]


def _stream(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def _gemm(A, a_strides, B, b_strides, Cm, M, N, K, *, beta=0.0, epi=_cabi.EPI_NONE, bias=None, save=None,
          saved=None, seed=0, uid=0, p=0.0, seed_ptr=None) -> None:
    g = _cabi.GemmArgs()
    g.A, g.B, g.C = A.data_ptr(), B.data_ptr(), Cm.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.as0, g.as1 = a_strides
    g.bs0, g.bs1 = b_strides
    g.cs0 = Cm.stride(0)
    g.beta, g.epilogue = beta, epi
    g.bias = bias.data_ptr() if bias is not None else None
    g.save = save.data_ptr() if save is not None else None
    g.saved = saved.data_ptr() if saved is not None else None
    g.seed, g.layer_uid, g.p_drop = seed, uid, p
    g.seed_ptr = seed_ptr
    dev = Cm.device
    _cabi.check(_cabi.lib().bcnf_train_gemm(C.byref(g), dev.index or 0, _stream(dev)), "bcnf_train_gemm")


def _colsum(X: torch.Tensor, out: torch.Tensor) -> None:
    dev = X.device
    _cabi.check(_cabi.lib().bcnf_train_colsum(X.data_ptr(), X.shape[0], X.shape[1], X.stride(0), out.data_ptr(), 0.0,
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


def _mlp_forward(u: torch.Tensor, ws: Sequence[torch.Tensor], bs: Sequence[torch.Tensor], spec: _Spec, li: int,
                 net: int):
    """u: (B, din + C) = cat([y_half, h]) (cnf.py:101).  Returns (o, saved activations)."""
    B = u.shape[0]
    acts, pres = [u], []
    for l in range(len(ws) - 1):
        n_out, n_in = ws[l].shape
        pre = torch.empty(B, n_out, device=u.device)
        out = torch.empty(B, n_out, device=u.device)
        x = acts[-1]
        # out = dropout(gelu(x W^T + b)): nn.Linear, nn.GELU, nn.Dropout (cnf.py:79-83)
        _gemm(x, (x.stride(0), 1), ws[l], (1, ws[l].stride(0)), out, B, n_out, n_in, epi=_cabi.EPI_BIAS_GELU_DROP,
              bias=bs[l], save=pre, seed=spec.seed, uid=layer_uid(li, net, l), p=spec.p_drop, seed_ptr=spec.seed_ptr)
        pres.append(pre)
        acts.append(out)
    n_out, n_in = ws[-1].shape
    o = torch.empty(B, n_out, device=u.device)
    x = acts[-1]
    _gemm(x, (x.stride(0), 1), ws[-1], (1, ws[-1].stride(0)), o, B, n_out, n_in, epi=_cabi.EPI_BIAS, bias=bs[-1])
    return o, acts, pres


def _mlp_backward(do: torch.Tensor, ws, bs, acts, pres, spec: _Spec, li: int, net: int):
    """Returns (du, [dW...], [db...]) for one conditioner network."""
    B = do.shape[0]
    dws, dbs = [None] * len(ws), [None] * len(ws)
    d_out = do.contiguous()
    for l in range(len(ws) - 1, -1, -1):
        n_out, n_in = ws[l].shape
        x = acts[l]
        # dW = d_out^T x   (A(i=n, r=m) = d_out[m, n];  B(r=m, j=k) = x[m, k])
        dw = torch.empty_like(ws[l])
        _gemm(d_out, (1, d_out.stride(0)), x, (x.stride(0), 1), dw, n_out, n_in, B)
        db = torch.empty_like(bs[l])
        _colsum(d_out, db)
        dws[l], dbs[l] = dw, db
        # d_in = d_out W  (* gelu'(pre) * mask of the layer below)
        d_in = torch.empty(B, n_in, device=do.device)
        if l > 0:
            _gemm(d_out, (d_out.stride(0), 1), ws[l], (ws[l].stride(0), 1), d_in, B, n_in, n_out,
                  epi=_cabi.EPI_DGELU_DROP, saved=pres[l - 1], seed=spec.seed, uid=layer_uid(li, net, l - 1), p=spec.p_drop,
                  seed_ptr=spec.seed_ptr)
        else:
            _gemm(d_out, (d_out.stride(0), 1), ws[l], (ws[l].stride(0), 1), d_in, B, n_in, n_out)
        d_out = d_in
    return d_out, dws, dbs


class _StackFn(torch.autograd.Function):
    """z, log|det J| = stack(y, h; parameters) with a hand-written backward."""

    @staticmethod
    def forward(ctx, spec: _Spec, y: torch.Tensor, h: torch.Tensor, *params: torch.Tensor):
        D = spec.size
        da = (D + 1) // 2
        y = y.contiguous().float()
        h = h.contiguous().float()
        B = y.shape[0]
        ld = torch.zeros(B, device=y.device)
        saved: list[Any] = []
        it = iter(params)
        nets = 2 if spec.two_way else 1
        for li, kind in enumerate(spec.kinds):
            if kind == "actnorm":
                scale, bias = next(it), next(it)
                saved.append((y,))
                y = scale * y + bias                                   # cnf.py:349
                ld = ld + torch.log(torch.abs(scale)).sum()            # cnf.py:350
            elif kind == "ortho":
                q = next(it)
                saved.append(())
                y = y @ q                                              # cnf.py:335
            else:
                per_net = []
                for net in range(nets):
                    ws = [next(it) for _ in range(spec.n_lin)]
                    bs = [next(it) for _ in range(spec.n_lin)]
                    src = slice(0, da) if net == 0 else slice(da, D)   # nn_a reads y_a, nn_b reads z_b (cnf.py:178, :183)
                    dst = slice(da, D) if net == 0 else slice(0, da)
                    u = torch.cat([y[:, src], h], dim=1)               # cnf.py:101
                    o, acts, pres = _mlp_forward(u, ws, bs, spec, li, net)
                    half = o.shape[1] // 2
                    t, ls = o[:, :half], torch.tanh(o[:, half:])       # cnf.py:104, :107
                    e = torch.exp(ls)
                    y_dst = y[:, dst]
                    new = e * y_dst + t                                # cnf.py:179 / :184
                    ld = ld + ls.sum(dim=1)                            # cnf.py:190, :193
                    y = torch.cat([y[:, :da], new], 1) if net == 0 else torch.cat([new, y[:, da:]], 1)
                    per_net.append((acts, pres, ls, e, y_dst))
                saved.append(tuple(per_net))
        ctx.spec, ctx.saved_acts, ctx.params = spec, saved, params
        ctx.h_cols = h.shape[1]
        return y, ld

    @staticmethod
    def backward(ctx, dz: torch.Tensor, dld: torch.Tensor):
        spec, saved, params = ctx.spec, ctx.saved_acts, ctx.params
        D = spec.size
        da = (D + 1) // 2
        nets = 2 if spec.two_way else 1
        dz = dz.contiguous().float()
        B = dz.shape[0]
        dld = dld.contiguous().float() if dld is not None else torch.zeros(B, device=dz.device)
        dh = torch.zeros(B, ctx.h_cols, device=dz.device)
        grads: list[Any] = [None] * len(params)
        # parameter offsets per layer, in forward order
        offs, o = [], 0
        for kind in spec.kinds:
            offs.append(o)
            o += 2 if kind == "actnorm" else 1 if kind == "ortho" else 2 * spec.n_lin * nets
        for li in range(len(spec.kinds) - 1, -1, -1):
            kind, po = spec.kinds[li], offs[li]
            if kind == "actnorm":
                scale = params[po]
                (x,) = saved[li]
                grads[po] = (dz * x).sum(0) + dld.sum() / scale        # d/ds [s x + b] and d/ds sum log|s|
                grads[po + 1] = dz.sum(0)
                dz = dz * scale
            elif kind == "ortho":
                dz = dz @ params[po].t()                               # Q is frozen (requires_grad False)
            else:
                for net in range(nets - 1, -1, -1):
                    acts, pres, ls, e, y_dst = saved[li][net]
                    base = po + net * 2 * spec.n_lin
                    ws = params[base: base + spec.n_lin]
                    bs = params[base + spec.n_lin: base + 2 * spec.n_lin]
                    src = slice(0, da) if net == 0 else slice(da, D)
                    dst = slice(da, D) if net == 0 else slice(0, da)
                    d_new = dz[:, dst]
                    dt = d_new
                    dls = d_new * y_dst * e + dld.unsqueeze(1)         # via z = e*y + t and via log-det
                    do = torch.cat([dt, dls * (1.0 - ls * ls)], dim=1)  # tanh'
                    du, dws, dbs = _mlp_backward(do, ws, bs, acts, pres, spec, li, net)
                    for j in range(spec.n_lin):
                        grads[base + j] = dws[j]
                        grads[base + spec.n_lin + j] = dbs[j]
                    d_src = dz[:, src] + du[:, : src.stop - src.start]
                    d_dst = d_new * e
                    dz = torch.cat([d_src, d_dst], 1) if net == 0 else torch.cat([d_dst, d_src], 1)
                    dh = dh + du[:, src.stop - src.start:]
        needs = ctx.needs_input_grad
        out_params = [g if needs[3 + i] else None for i, g in enumerate(grads)]
        return (None, dz if needs[1] else None, dh if needs[2] else None, *out_params)


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
    capturable optimizer (``torch.optim.Adam(..., capturable=True)``) and a single process (no DDP).
    """

    def __init__(self, model: Any, optimizer: torch.optim.Optimizer, hybrid_weight: float = 0.0,
                 cuda_graph: bool = False) -> None:
        from .utils import inn_nll_loss
        self.model, self.optimizer, self.hybrid_weight = model, optimizer, hybrid_weight
        self.loss_function = inn_nll_loss
        self.mse_loss = torch.nn.MSELoss()
        self.cuda_graph = cuda_graph
        self._graph: Any = None
        self._static: Any = None
        if cuda_graph:
            if hasattr(model, "module"):
                raise NotImplementedError("cuda_graph=True with DistributedDataParallel is not supported")
            if not all(g.get("capturable", False) for g in optimizer.param_groups):
                raise ValueError("cuda_graph=True needs a capturable optimizer, e.g. torch.optim.Adam(..., capturable=True)")

    def _net(self) -> Any:
        return self.model.module if hasattr(self.model, "module") else self.model

    def _losses(self, y: torch.Tensor, *conditions: torch.Tensor):
        net = self._net()
        dev = net.device
        z, h = self.model(y.to(dev), *[c.to(dev) for c in conditions], log_det_J=True, return_features=True)
        if self.hybrid_weight > 0:
            mse = self.mse_loss(net.prediction_head(h), y.to(dev))
        else:
            mse = torch.zeros((), device=z.device)
        nll = self.loss_function(z, net.log_det_J)
        loss = (nll + mse * self.hybrid_weight) / (1 + self.hybrid_weight)
        return loss, nll, mse, z

    def _graphed_step(self, y: torch.Tensor, *conditions: torch.Tensor):
        net = self._net()
        dev = torch.device(net.device)
        shapes = (tuple(y.shape),) + tuple(tuple(c.shape) for c in conditions)
        if self._graph is None or self._static["shapes"] != shapes:
            net._dropout_seed_word = torch.zeros(1, dtype=torch.int64, device=dev)
            st = {"shapes": shapes, "y": torch.empty_like(y, device=dev),
                  "c": [torch.empty_like(c, device=dev) for c in conditions]}
            st["y"].copy_(y)
            for d, c in zip(st["c"], conditions):
                d.copy_(c)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):              # warm-up off the capture stream (allocator, lazy inits)
                for _ in range(3):
                    self.optimizer.zero_grad(set_to_none=True)
                    loss, _, _, _ = self._losses(st["y"], *st["c"])
                    loss.backward()
                    self.optimizer.step()
                    net._dropout_seed_word.add_(1)
            torch.cuda.current_stream(dev).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                loss, nll, mse, _ = self._losses(st["y"], *st["c"])
                loss.backward()
                self.optimizer.step()
                net._dropout_seed_word.add_(1)          # fresh dropout masks on every replay
            st.update(loss=loss, nll=nll, mse=mse)
            self._graph, self._static = graph, st
        st = self._static
        st["y"].copy_(y, non_blocking=True)
        for d, c in zip(st["c"], conditions):
            d.copy_(c, non_blocking=True)
        self._graph.replay()
        return st["loss"], st["nll"], st["mse"]

    def train_batch(self, y: torch.Tensor, *conditions: torch.Tensor) -> tuple[float, float, float]:
        if self.cuda_graph:
            loss, nll, mse = self._graphed_step(y, *conditions)
            return loss.item(), nll.item(), mse.item()
        self.optimizer.zero_grad()
        loss, nll, mse, _ = self._losses(y, *conditions)
        loss.backward()
        self.optimizer.step()
        torch.nn.utils.clip_grad_norm_(self._net().parameters(), max_norm=1.0)
        return loss.item(), nll.item(), mse.item()

    def validate_batch(self, y: torch.Tensor, *conditions: torch.Tensor):
        with torch.no_grad():
            loss, nll, mse, z = self._losses(y, *conditions)
        return loss.item(), nll.item(), mse.item(), z.mean(dim=0), z.std(dim=0)
