"""Condition encoders (adjacent to the hot path; plain PyTorch modules whose eval-mode forward on a tensor-core handle
is routed to bcnf_b200/feature_tc.py).

Same class names, constructor arguments and parameter names as the reference's
``src/bcnf/models/feature_network.py`` so that ``state_dict`` keys under
``feature_network_stack.feature_networks.{i}.`` match.  The B200 build changes only how often
they run: once per conditioning instance, never once per (sample, instance) row
(north star; the reference re-runs them on tiled conditions, cnf.py:579 + :497).
"""
from __future__ import annotations

import math
import warnings
from typing import Any, Type

import torch
import torch.nn.functional as F
from torch import nn

__all__ = ["FeatureNetwork", "FeatureNetworkStack", "ConcatenateCondition", "FrExpFeatureNetwork",
           "FullyConnectedFeatureNetwork", "LSTMFeatureNetwork", "MultiHeadAttention", "TransformerBlock",
           "Transformer"]


def _inference_call(x: torch.Tensor) -> bool:
    """Eval-mode calls take the tensor-core path unless the caller is differentiating with respect to the INPUT.

    The kernels return tensors without autograd history.  In eval mode the coupling stack itself runs without history
    (CondRealNVP_v2.forward), so a history on h would reach nothing; keying on ``torch.is_grad_enabled()`` alone made
    ``model.log_prob(...)`` called outside ``torch.no_grad()`` seven times slower than inside it.
    """
    return not (torch.is_grad_enabled() and x.requires_grad)


# ---- training: parameter gradients of the encoder's Linears off the dependency chain ------------------------------------
# Set by bcnf_b200.train.Trainer for the duration of a step's forward pass (None otherwise: plain nn.Linear autograd).
# At batch 256 the encoder's backward is a chain of ~500 launch-bound kernels; a Linear's weight gradient g^T x and bias
# gradient sum(g) feed nothing but the optimizer, so they do not have to sit on that chain: the Function below returns
# only dx and leaves the parameter gradients to a side stream (OffChain.defer), which the Trainer joins before the
# gradient all-reduce / optimizer step.
_OFF_CHAIN: Any = None


def _wgrad(g2: torch.Tensor, x2: torch.Tensor) -> torch.Tensor:
    """g2^T x2 for tall operands (rows >> columns).  cuBLAS runs the plain product as a handful of CTAs over the whole row
    count (42-114 us for 7680 x 128: measured, profiles/r02b_train_step_kernels_cupti.txt); cut into row slabs it is one
    batched GEMM that fills the GPU plus a small sum."""
    rows = g2.shape[0]
    for slabs in (32, 16, 8, 4):
        if rows % slabs == 0 and rows // slabs >= 128 and g2.is_contiguous() and x2.is_contiguous():
            r = rows // slabs
            return torch.bmm(g2.view(slabs, r, -1).transpose(1, 2), x2.view(slabs, r, -1)).sum(0)
    return g2.t().mm(x2)


class OffChain:
    """Side streams for the deferred parameter gradients of one training step."""

    def __init__(self, streams: list) -> None:
        self.streams, self.keep, self.n = streams, [], 0

    @staticmethod
    def accumulate(p: torch.Tensor | None, d: torch.Tensor) -> None:
        """p.grad += d, as autograd's AccumulateGrad would (called on the side stream that computed d)."""
        if p is None or not p.requires_grad:
            return
        if p.grad is None:
            p.grad = d
        else:
            p.grad.add_(d)

    def run(self, fn: Any, *keep: Any) -> None:
        """fn() on the next side stream, after everything enqueued so far on the current stream.  ``keep``: the tensors fn
        reads -- held until join(), so that the caching allocator cannot hand their memory out while fn is pending."""
        st = self.streams[self.n % len(self.streams)]
        main = torch.cuda.current_stream(st.device)
        self.n += 1
        st.wait_stream(main)
        with torch.cuda.stream(st):
            fn()
        self.keep.append(keep)

    def defer(self, x: torch.Tensor, g: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None) -> None:
        """Parameter gradients of y = x W^T + b given g = dL/dy: dW = g^T x, db = sum over rows of g."""
        g2, x2 = g.reshape(-1, g.shape[-1]), x.reshape(-1, x.shape[-1])

        def grads():
            self.accumulate(weight, _wgrad(g2, x2))
            if bias is not None:
                self.accumulate(bias, g2.sum(0))
        self.run(grads, g2, x2)

    def join(self, main: Any) -> None:
        if self.n:
            for st in self.streams:
                main.wait_stream(st)
        self.keep.clear()
        self.n = 0


class _OffChainLinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None, oc: OffChain):
        ctx.save_for_backward(x, weight)
        ctx.bias, ctx.oc = bias, oc
        return F.linear(x, weight, bias)

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        x, weight = ctx.saved_tensors
        dx = g.matmul(weight) if ctx.needs_input_grad[0] else None
        ctx.oc.defer(x, g, weight, ctx.bias)
        return dx, None, None, None


class _OffChainLayerNormFn(torch.autograd.Function):
    """nn.LayerNorm whose backward computes dx on the chain and d gamma / d beta (a column reduction over all token rows
    that only the optimizer reads) on a side stream: the same ATen kernels, selected through their output mask."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float, oc: OffChain):
        out, mean, rstd = torch.ops.aten.native_layer_norm(x, [x.shape[-1]], weight, bias, eps)
        ctx.save_for_backward(x, mean, rstd, weight)
        ctx.bias, ctx.oc = bias, oc
        return out

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        x, mean, rstd, weight = ctx.saved_tensors
        g = g.contiguous()
        shape = [x.shape[-1]]
        dx = torch.ops.aten.native_layer_norm_backward(g, x, shape, mean, rstd, weight, ctx.bias, [True, False, False])[0]
        oc, bias = ctx.oc, ctx.bias

        def params_grad():
            _, dw, db = torch.ops.aten.native_layer_norm_backward(g, x, shape, mean, rstd, weight, bias, [False, True, True])
            oc.accumulate(weight, dw)
            oc.accumulate(bias, db)
        oc.run(params_grad, g, x, mean, rstd)
        return dx, None, None, None, None


def _layer_norm(ln: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    oc = _OFF_CHAIN
    if oc is not None and x.is_cuda and torch.is_grad_enabled() and ln.elementwise_affine and ln.bias is not None \
            and ln.weight.requires_grad and len(ln.normalized_shape) == 1:
        return _OffChainLayerNormFn.apply(x, ln.weight, ln.bias, ln.eps, oc)
    return ln(x)


def _linear(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """lin(x); inside a Trainer step on a CUDA device with the parameter gradients deferred to a side stream."""
    oc = _OFF_CHAIN
    if oc is not None and x.is_cuda and torch.is_grad_enabled() and lin.weight.requires_grad:
        return _OffChainLinearFn.apply(x, lin.weight, lin.bias, oc)
    return lin(x)


class FeatureNetwork(nn.Module):
    """Base class: records ``input_size`` / ``output_size`` (reference feature_network.py:10-25)."""
    input_size: int
    output_size: int

    @property
    def n_params(self) -> int:
        return sum(p.numel() for p in self.parameters())


class ConcatenateCondition(FeatureNetwork):
    """Marker that consumes one raw condition (reference feature_network.py:76-88)."""

    def __init__(self, input_size: int, output_size: int, dim: int = -1) -> None:
        super().__init__()
        self.input_size, self.output_size, self.dim = input_size, output_size, dim

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x


class FeatureNetworkStack(FeatureNetwork):
    """Chain of feature networks; each ConcatenateCondition consumes the next raw condition
    (reference feature_network.py:28-73)."""

    def __init__(self, feature_networks: list[FeatureNetwork | nn.Module | None] | None = None) -> None:
        super().__init__()
        nets = [fn for fn in (feature_networks or []) if fn is not None]
        if not nets:
            raise ValueError("Feature network stack must contain at least one feature network.")
        self.feature_networks = nn.Sequential(*nets)
        self.n_distinct_conditions = sum(isinstance(fn, ConcatenateCondition) for fn in nets)
        self.input_size = getattr(nets[0], "input_size", None)
        self.output_size = getattr(nets[-1], "output_size", None)

    def forward(self, *conditions: torch.Tensor, skip_last: bool = False) -> torch.Tensor:
        """``skip_last`` (bcnf_b200 only): stop in front of the last network and return its input -- the caller fuses
        that network's affine output layer with the condition projection (feature_tc.fused_projection)."""
        if len(conditions) != self.n_distinct_conditions:
            raise ValueError(f"Expected {self.n_distinct_conditions} conditions, but got {len(conditions)}.")
        taken = 0
        feats: torch.Tensor | None = None
        nets = list(self.feature_networks)
        for fn in nets[:-1] if skip_last else nets:
            if isinstance(fn, ConcatenateCondition):
                raw = conditions[taken]
                taken += 1
                feats = fn(raw if feats is None else torch.cat([feats, raw], dim=fn.dim))
            else:
                feats = fn(feats)
        return feats


class FrExpFeatureNetwork(FeatureNetwork):
    """Mantissa / exponent split of the input (reference feature_network.py:91-111)."""

    def __init__(self, input_size: int, separate_sign: bool = False) -> None:
        super().__init__()
        self.separate_sign = separate_sign
        self.input_size = input_size
        self.output_size = input_size * (3 if separate_sign else 2)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        mant, expo = torch.frexp(x)
        if self.separate_sign:
            return torch.cat([torch.sign(mant), mant.abs(), expo], dim=-1)
        return torch.cat([mant, expo], dim=-1)


class FullyConnectedFeatureNetwork(FeatureNetwork):
    """Flatten + MLP (reference feature_network.py:114-145); parameters live in ``self.nn``."""

    def __init__(self, sizes: list[int], activation: Type[nn.Module] = nn.GELU, dropout: float = 0.0,
                 batch_norm: bool = False) -> None:
        super().__init__()
        self.input_size, self.output_size = sizes[0], sizes[-1]
        self.output_size_lin = sizes[-1]
        self.nn = nn.Sequential()
        if len(sizes) < 2:
            warnings.warn("No hidden layers in the fully connected network. Using identity function.")
            self.nn.append(nn.Identity())
            return
        for fan_in, fan_out in zip(sizes[:-2], sizes[1:-1]):
            self.nn.append(nn.Linear(fan_in, fan_out))
            if batch_norm:
                self.nn.append(nn.BatchNorm1d(fan_out))
            self.nn.append(activation())
            if dropout > 0.0:
                self.nn.append(nn.Dropout(dropout))
        self.nn.append(nn.Linear(sizes[-2], sizes[-1]))

    tc_passes: int = 0      # set by CondRealNVP_v2 on tensor-core handles: 3 = bf16x3 (fp32-class), 1 = bf16, 0 = PyTorch

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.tc_passes and x.is_cuda and not self.training and _inference_call(x):
            from . import feature_tc
            if x.size(0) >= feature_tc.MIN_ROWS and feature_tc.supported(self):
                return feature_tc.forward(self, x, self.tc_passes)
        return self.nn(x.reshape(x.size(0), -1))


class LSTMFeatureNetwork(FeatureNetwork):
    """(bi)LSTM -> Linear -> pooling (reference feature_network.py:148-178).

    Deviation, documented in DESIGN.md: the reference builds the LSTM with ``batch_first=True``
    but pools with ``dim=0``, i.e. over the BATCH axis, so its output is (seq_len, output_size)
    and the flow only runs when the batch size equals the sequence length (SURVEY.md section 8a
    hazard).  ``pool_axis="time"`` (default) pools over the sequence axis, giving one feature
    vector per instance, which is what every caller assumes; ``pool_axis="reference"``
    reproduces the reference exactly.
    """

    def __init__(self, input_size: int, hidden_size: int, output_size: int, num_layers: int, dropout: float = 0.0,
                 bidirectional: bool = False, pooling: str = "mean", pool_axis: str = "time") -> None:
        super().__init__()
        if pooling not in ("mean", "max"):
            raise ValueError(f'Pooling method {pooling} not supported. Use either "mean" or "max".')
        if pool_axis not in ("time", "reference"):
            raise ValueError("pool_axis must be 'time' or 'reference'")
        self.input_size, self.output_size = input_size, output_size
        self.pooling, self.pool_axis = pooling, pool_axis
        self.lstm = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers, dropout=dropout,
                            bidirectional=bidirectional, batch_first=True)
        self.linear = nn.Linear(hidden_size * (2 if bidirectional else 1), output_size)

    tc_passes: int = 0      # set by CondRealNVP_v2 on tensor-core handles: 3 = bf16x3 (fp32-class), 1 = bf16, 0 = PyTorch

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.tc_passes and x.is_cuda and x.ndim == 3 and not self.training and _inference_call(x):
            from . import feature_tc
            if x.size(0) >= feature_tc.MIN_ROWS_LSTM and feature_tc.lstm_supported(self):
                return feature_tc.lstm_forward(self, x, self.tc_passes)
        seq, _ = self.lstm(x)
        axis = 1 if self.pool_axis == "time" else 0
        if self.pooling == "mean":
            # mean pooling commutes with the affine output layer: pool first, then ONE Linear per pooled row instead of
            # one per time step (reference feature_network.py:168-176 applies the Linear to all seq_len x batch rows:
            # 30x the FLOPs and a (B, T, output_size) intermediate; same result up to fp32 rounding)
            return self.linear(seq.mean(dim=axis))
        return self.linear(seq).max(dim=axis).values


class MultiHeadAttention(nn.Module):
    """Reference feature_network.py:183-229 (parameter names q_linear/k_linear/v_linear/fc_out)."""

    def __init__(self, d_model: int, n_heads: int) -> None:
        super().__init__()
        self.d_model, self.n_heads, self.head_dim = d_model, n_heads, d_model // n_heads
        self.q_linear = nn.Linear(d_model, d_model)
        self.k_linear = nn.Linear(d_model, d_model)
        self.v_linear = nn.Linear(d_model, d_model)
        self.fc_out = nn.Linear(d_model, d_model)

    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor, mask: Any = None) -> torch.Tensor:
        b = query.size(0)

        def heads(t: torch.Tensor, lin: nn.Linear) -> torch.Tensor:
            return _linear(lin, t).view(b, -1, self.n_heads, self.head_dim).transpose(1, 2)

        q, k, v = heads(query, self.q_linear), heads(key, self.k_linear), heads(value, self.v_linear)
        scores = q @ k.transpose(-2, -1) / math.sqrt(self.head_dim)
        if mask is not None:
            scores = scores.masked_fill(mask == 0, -1e9)
        ctx = F.softmax(scores, dim=-1) @ v
        return _linear(self.fc_out, ctx.transpose(1, 2).contiguous().view(b, -1, self.d_model))


class TransformerBlock(nn.Module):
    """Post-norm block (reference feature_network.py:232-260)."""

    def __init__(self, d_model: int, n_heads: int, ff_size: int, dropout: float = 0.1) -> None:
        super().__init__()
        self.attention = MultiHeadAttention(d_model, n_heads)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.ffn = nn.Sequential(nn.Linear(d_model, ff_size), nn.GELU(), nn.Linear(ff_size, d_model))
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = _layer_norm(self.norm1, x + self.dropout(self.attention(x, x, x)))
        if _OFF_CHAIN is not None and len(self.ffn) == 3:
            f = _linear(self.ffn[2], self.ffn[1](_linear(self.ffn[0], x)))
        else:
            f = self.ffn(x)
        return _layer_norm(self.norm2, x + self.dropout(f))


class Transformer(FeatureNetwork):
    """Token embedding + blocks + Linear on token 0 (reference feature_network.py:263-307)."""

    def __init__(self, input_size: int, trf_size: int, n_heads: int, ff_size: int, n_blocks: int, output_size: int,
                 dropout: float = 0.5, trf_dropout: float = 0.1, add_positional_embeddings: bool = False) -> None:
        super().__init__()
        self.input_size, self.output_size = input_size, output_size
        self.add_positional_embeddings = add_positional_embeddings
        self.trf_size = trf_size
        self.features = nn.Linear(input_size, trf_size)
        self.layers = nn.ModuleList([TransformerBlock(trf_size, n_heads, ff_size, trf_dropout) for _ in range(n_blocks)])
        self.output = nn.Linear(trf_size, output_size)
        self.dropout = nn.Dropout(dropout)

    def _positional(self, seq_len: int, device: torch.device) -> torch.Tensor:
        # only the first `input_size` channels are filled, as in the reference (:293-299)
        pe = torch.zeros(seq_len, self.trf_size, device=device)
        pos = torch.arange(seq_len, dtype=torch.float64).unsqueeze(1)
        j = torch.arange(self.input_size, dtype=torch.float64).unsqueeze(0)
        ang = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * j / self.input_size)
        vals = torch.where((torch.arange(self.input_size) % 2 == 0).unsqueeze(0), torch.sin(ang), torch.cos(ang))
        pe[:, : self.input_size] = vals.to(torch.float32).to(device)
        return pe

    tc_passes: int = 0      # set by CondRealNVP_v2 on tensor-core handles: 3 = bf16x3 (fp32-class), 1 = bf16, 0 = PyTorch

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.tc_passes and x.is_cuda and x.ndim == 3 and not self.training and _inference_call(x):
            from . import feature_tc
            if x.size(0) >= feature_tc.MIN_ROWS_TRF and feature_tc.transformer_supported(self, x.size(1)):
                return feature_tc.transformer_forward(self, x, self.tc_passes)
        if _OFF_CHAIN is not None and self.training and x.is_cuda and torch.is_grad_enabled():
            from . import trf_train            # inside a Trainer step: own kernels forward, hand-written backward
            if trf_train.usable(self, x):
                return trf_train.forward(self, x, _OFF_CHAIN)
        x = self.dropout(_linear(self.features, x))
        if self.add_positional_embeddings:
            x = x + self._positional(x.size(1), x.device)
        for layer in self.layers:
            x = layer(x)
        x = self.dropout(x)
        return _linear(self.output, x[:, 0, :])
