"""CondRealNVP_v2 on B200: the reference's Python surface over the bcnf_b200 C ABI.

Drop-in for ``bcnf.models.cnf`` of psaegert/bcnf (reference ``src/bcnf/models/cnf.py``): same
class names, constructor arguments, methods, side effects (``model.log_det_J``) and, above
all, the same ``state_dict`` keys, so checkpoints and the ``configs/runs/*.yaml`` files load
unchanged.  The modules below are parameter containers; every flow computation -- the whole
stack, or one layer when a layer is called on its own -- is one call into
``libbcnf_b200.so`` (include/bcnf_b200.h).  There is no PyTorch or CPU fallback.

Differences from the reference are listed in DESIGN.md ("Deviations").
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Any, Sequence

import torch
import torch.nn as nn

from . import _cabi
from .factories import FeatureNetworkFactory
from .feature_network import FeatureNetwork, FeatureNetworkStack
from .utils import ParameterIndexMapping

__all__ = ["InvertibleLayer", "ConditionalInvertibleLayer", "ConditionalNestedNeuralNetwork",
           "ConditionalAffineCouplingLayer", "OrthonormalTransformation", "ActNorm", "CondRealNVP_v2"]


def _as_device(device: Any) -> torch.device:
    return device if isinstance(device, torch.device) else torch.device(device)


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _dev_f32(t: torch.Tensor, device: torch.device, what: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise NotImplementedError(f"{what}: dtype {t.dtype} -- the B200 path computes in float32 "
                                  "(every reference config sets global.dtype float32)")
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


PRECISIONS = {"auto": -1, "fp32": _cabi.PREC_FP32, "bf16x3": _cabi.PREC_BF16X3, "bf16": _cabi.PREC_BF16}


# ------------------------------------------------------------------------------------------
# packed handle
# ------------------------------------------------------------------------------------------
class PackedFlow:
    """A ``bcnf_flow_t`` built from a list of layer modules; repacks when parameters change."""
    check_row_map = True     # validate explicit row2inst indices against P (one device reduction + sync per call)

    def __init__(self, layers: Sequence[nn.Module], size: int, n_conditions: int, nested_sizes: Sequence[int],
                 two_way: bool, device: torch.device, precision: str = "fp32") -> None:
        if device.type != "cuda":
            raise RuntimeError(f"bcnf_b200 runs on CUDA devices only (got {device}); there is no CPU path. "
                               "Move the model with .to('cuda').")
        self.lib = _cabi.lib()
        self.layers = list(layers)
        self.device = device
        self.size = size
        types = []
        for layer in self.layers:
            if isinstance(layer, ActNorm):
                types.append(_cabi.OP_ACTNORM)
            elif isinstance(layer, ConditionalAffineCouplingLayer):
                types.append(_cabi.OP_COUPLING)
            elif isinstance(layer, OrthonormalTransformation):
                types.append(_cabi.OP_ORTHO)
            else:   # same complaint as the reference's layer loop, cnf.py:485
                raise ValueError("Layer must be an instance of ConditionalInvertibleLayer or InvertibleLayer, "
                                 f"but got {type(layer)}")
        self.types = types
        desc = _cabi.FlowDesc()
        desc.size, desc.n_conditions, desc.n_hidden = size, n_conditions, len(nested_sizes)
        if len(nested_sizes) > _cabi.MAX_HIDDEN_LAYERS:
            raise NotImplementedError(f"len(nested_sizes)={len(nested_sizes)} > {_cabi.MAX_HIDDEN_LAYERS}")
        for i, h in enumerate(nested_sizes):
            desc.hidden[i] = int(h)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        desc.two_way, desc.n_ops = int(two_way), len(types)
        desc.device = device.index if device.index is not None else torch.cuda.current_device()
        self._handle = C.c_void_p()
        arr = (C.c_int32 * len(types))(*types)
        has_coupling = _cabi.OP_COUPLING in types
        if precision == "auto":
            # tensor-core 3-pass split (fp32-class accuracy) where the conditioner is wide enough to be a
            # real GEMM, fp32 FMA kernels otherwise -- a choice between CUDA kernels, not a fallback
            wide = has_coupling and min(nested_sizes) >= 48
            precision = "bf16x3" if wide else "fp32"
            desc.precision = PRECISIONS[precision]
            rc = self.lib.bcnf_flow_create(C.byref(desc), arr, C.byref(self._handle))
            if rc == -2 and precision != "fp32":
                precision = "fp32"
                desc.precision = PRECISIONS[precision]
                rc = self.lib.bcnf_flow_create(C.byref(desc), arr, C.byref(self._handle))
            _cabi.check(rc, "bcnf_flow_create")
        else:
            desc.precision = PRECISIONS[precision] if has_coupling else _cabi.PREC_FP32
            _cabi.check(self.lib.bcnf_flow_create(C.byref(desc), arr, C.byref(self._handle)), "bcnf_flow_create")
        self.precision = precision
        info = _cabi.FlowInfo()
        _cabi.check(self.lib.bcnf_flow_info(self._handle, C.byref(info)), "bcnf_flow_info")
        self.info = info
        self.proj_width = int(info.proj_width)
        self.kernel = _cabi.KERNEL_NAMES[int(info.kernel)]
        self._sig = None

    def __del__(self) -> None:
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self.lib.bcnf_flow_destroy(h)
            except Exception:
                pass
            self._handle = C.c_void_p()

    # -- parameters ---------------------------------------------------------------------
    def _params(self) -> list[torch.Tensor]:
        out: list[torch.Tensor] = []
        for layer in self.layers:
            out.extend(p for p in layer.parameters())
        return out

    def sync_params(self) -> None:
        sig = tuple((p.data_ptr(), p._version) for p in self._params())
        if sig == self._sig:
            return
        keep: list[Any] = []

        def ptr(t: torch.Tensor) -> int:
            t = _dev_f32(t.detach(), self.device, "parameter")
            keep.append(t)
            return t.data_ptr()

        def ptr_array(ts: list[torch.Tensor]):
            a = (C.c_void_p * len(ts))(*[ptr(t) for t in ts])
            keep.append(a)
            return C.cast(a, _cabi.FloatPP)

        ops = (_cabi.OpParams * len(self.layers))()
        for i, (layer, ty) in enumerate(zip(self.layers, self.types)):
            ops[i].type = ty
            if ty == _cabi.OP_ACTNORM:
                ops[i].scale, ops[i].bias = ptr(layer.scale), ptr(layer.bias)
            elif ty == _cabi.OP_ORTHO:
                ops[i].q = ptr(layer.orthonormal_matrix)
            else:
                lin = layer.nn_a.linears()
                ops[i].w_a = ptr_array([m.weight for m in lin])
                ops[i].b_a = ptr_array([m.bias for m in lin])
                if layer.two_way:
                    lin = layer.nn_b.linears()
                    ops[i].w_b = ptr_array([m.weight for m in lin])
                    ops[i].b_b = ptr_array([m.bias for m in lin])
        _cabi.check(self.lib.bcnf_flow_set_params(self._handle, ops, _stream_ptr(self.device)),
                    "bcnf_flow_set_params")
        self._sig = sig

    # -- compute ------------------------------------------------------------------------
    def project(self, h: torch.Tensor) -> torch.Tensor:
        """(n_inst, C) features -> (n_inst, proj_width) hoisted first-layer terms."""
        self.sync_params()
        h = _dev_f32(h, self.device, "features")
        p = torch.empty((h.shape[0], max(self.proj_width, 1)), dtype=torch.float32, device=self.device)
        if self.proj_width > 0:
            _cabi.check(self.lib.bcnf_cond_project(self._handle, h.data_ptr(), h.shape[0], p.data_ptr(),
                                                   _stream_ptr(self.device)), "bcnf_cond_project")
        return p

    def run(self, inverse: bool, x: torch.Tensor, P: torch.Tensor, *, row2inst: torch.Tensor | None = None,
            inst_period: int = 0, want_logdet: bool = False,
            out: torch.Tensor | None = None) -> tuple[torch.Tensor, torch.Tensor | None]:
        self.sync_params()
        x = _dev_f32(x, self.device, "input")
        if x.ndim != 2 or x.shape[1] != self.size:
            raise ValueError(f"expected input of shape (B, {self.size}), got {tuple(x.shape)}")
        n = x.shape[0]
        r2i = self._check_map(P, n, row2inst, inst_period)
        if out is None:
            out = torch.empty_like(x)
        ld = torch.empty(n, dtype=torch.float32, device=self.device) if want_logdet else None
        fn = self.lib.bcnf_flow_inverse if inverse else self.lib.bcnf_flow_forward
        _cabi.check(fn(self._handle, x.data_ptr(), P.data_ptr(), r2i, int(inst_period), n, out.data_ptr(),
                       ld.data_ptr() if ld is not None else 0, _stream_ptr(self.device)),
                    "bcnf_flow_inverse" if inverse else "bcnf_flow_forward")
        return out, ld


    def sample(self, n_rows: int, P: torch.Tensor, *, seed: int, sigma: float = 1.0, row2inst: torch.Tensor | None = None,
               inst_period: int = 0, out: torch.Tensor | None = None) -> torch.Tensor:
        """Inverse pass on a latent drawn INSIDE the kernel (``bcnf_flow_sample``): z = sigma * N(0, 1) from
        Philox4x32-10 at counter row * D + j under ``seed``.  No z array is allocated, written or read."""
        self.sync_params()
        if out is None:
            out = torch.empty((n_rows, self.size), dtype=torch.float32, device=self.device)
        r2i = self._check_map(P, n_rows, row2inst, inst_period)
        _cabi.check(self.lib.bcnf_flow_sample(self._handle, int(seed) & (2 ** 64 - 1), float(sigma), P.data_ptr(), r2i,
                                              int(inst_period), n_rows, out.data_ptr(), 0, _stream_ptr(self.device)),
                    "bcnf_flow_sample")
        return out

    def sample_ranks(self, n_rows: int, P: torch.Tensor, y: torch.Tensor, ranks: torch.Tensor, *, seed: int = 0,
                     sigma: float = 1.0, z: torch.Tensor | None = None, row2inst: torch.Tensor | None = None,
                     inst_period: int = 0) -> torch.Tensor:
        """``ranks[i, j] += #{rows of instance i : x[row, j] < y[i, j]}`` (``bcnf_flow_sample_ranks``): the reduction of
        calibration.py:44-48 in the kernel's output stage; the samples are never written.  ``z`` given: read instead of drawn."""
        self.sync_params()
        y = _dev_f32(y, self.device, "y")
        if y.shape != (P.shape[0], self.size) or ranks.shape != y.shape or ranks.dtype != torch.int32 or not ranks.is_contiguous():
            raise ValueError(f"y and ranks must be (n_inst, {self.size}) = {(P.shape[0], self.size)}, ranks int32 contiguous")
        zp = 0
        if z is not None:
            z = _dev_f32(z, self.device, "z")
            if z.shape != (n_rows, self.size):
                raise ValueError(f"expected z of shape ({n_rows}, {self.size}), got {tuple(z.shape)}")
            zp = z.data_ptr()
        r2i = self._check_map(P, n_rows, row2inst, inst_period)
        _cabi.check(self.lib.bcnf_flow_sample_ranks(self._handle, zp, int(seed) & (2 ** 64 - 1), float(sigma), P.data_ptr(), r2i,
                                                    int(inst_period), n_rows, y.data_ptr(), ranks.data_ptr(),
                                                    _stream_ptr(self.device)), "bcnf_flow_sample_ranks")
        return ranks

    def _check_map(self, P: torch.Tensor, n: int, row2inst: torch.Tensor | None, inst_period: int) -> int:
        """The ABI takes raw pointers: the row -> instance map is checked here, against the projection it indexes."""
        n_inst = P.shape[0]
        if self.proj_width == 0:        # no conditioner network in this (sub-)stack: P is a placeholder nobody reads
            return 0
        if P.ndim != 2 or P.shape[1] != max(self.proj_width, 1) or P.dtype != torch.float32 or not P.is_contiguous():
            raise ValueError(f"P must be a contiguous float32 (n_inst, {max(self.proj_width, 1)}) tensor, got "
                             f"{tuple(P.shape)} {P.dtype}")
        if row2inst is not None:
            if row2inst.numel() != n:
                raise ValueError(f"row2inst has {row2inst.numel()} entries for {n} rows")
            if n and self.check_row_map:
                lo, hi = int(row2inst.min()), int(row2inst.max())
                if lo < 0 or hi >= n_inst:
                    raise IndexError(f"row2inst entries must lie in [0, {n_inst}), got [{lo}, {hi}]")
            self._r2i_keep = row2inst.to(device=self.device, dtype=torch.int32).contiguous()
            return self._r2i_keep.data_ptr()
        if inst_period > 0:
            if inst_period > n_inst:
                raise IndexError(f"inst_period={inst_period} exceeds the {n_inst} instances of P")
        elif n > n_inst:
            raise IndexError(f"{n} rows but P holds {n_inst} instances (identity row -> instance map)")
        return 0


# ------------------------------------------------------------------------------------------
# layer modules (parameter containers with the reference's names and state_dict keys)
# ------------------------------------------------------------------------------------------
class InvertibleLayer(nn.Module):
    """Reference cnf.py:14-28."""
    log_det_J: float | torch.Tensor | None
    n_conditions: int

    @property
    def n_params(self) -> int:
        return sum(p.numel() for p in self.parameters())


class ConditionalInvertibleLayer(nn.Module):
    """Reference cnf.py:31-46."""
    log_det_J: float | torch.Tensor | None
    n_conditions: int

    @property
    def n_params(self) -> int:
        return sum(p.numel() for p in self.parameters())


def _require_linear_gelu(layer: str, layer_kwargs: Any, activation: str, activation_kwargs: Any) -> None:
    # The fused kernels implement nn.Linear + exact-erf nn.GELU (81 of the reference's 84 run
    # configs).  Anything else is refused rather than silently mis-computed.
    if layer != "Linear" or layer_kwargs:
        raise NotImplementedError(f"conditioner layer {layer!r} with kwargs {layer_kwargs!r}: only 'Linear' "
                                  "is implemented in the fused B200 kernels")
    if activation != "GELU" or activation_kwargs:
        raise NotImplementedError(f"conditioner activation {activation!r} with kwargs {activation_kwargs!r}: "
                                  "only 'GELU' (exact erf) is implemented in the fused B200 kernels")


class ConditionalNestedNeuralNetwork(nn.Module):
    """Parameters of the conditioner MLP (reference cnf.py:49-107).

    ``self.nn`` reproduces the reference's Sequential slot numbering -- Linear, GELU and, when
    ``dropout > 0``, Dropout per hidden layer (cnf.py:78-83) -- so ``nn.{j}.weight`` keys match.
    """

    def __init__(self, sizes: list[int], n_conditions: int, n_output_parameters: int, layer: str = "Linear",
                 layer_kwargs: dict[str, Any] | None = None, activation: str = "GELU",
                 activation_kwargs: dict[str, Any] | None = None, dropout: float = 0.0,
                 device: Any = "cpu") -> None:
        super().__init__()
        _require_linear_gelu(layer, layer_kwargs, activation, activation_kwargs)
        if len(sizes) < 3:
            raise NotImplementedError("a conditioner needs at least one hidden layer (nested_sizes must not be empty)")
        self.n_conditions = n_conditions
        self.n_output_parameters = n_output_parameters
        self.device = device
        widths = list(sizes)
        widths[0] = widths[0] + n_conditions             # cnf.py:72
        widths[-1] = widths[-1] * n_output_parameters    # cnf.py:75
        self.widths = widths
        self.nn = nn.Sequential()
        for fan_in, fan_out in zip(widths[:-2], widths[1:-1]):
            self.nn.append(nn.Linear(fan_in, fan_out))
            self.nn.append(nn.GELU())
            if dropout > 0.0:
                self.nn.append(nn.Dropout(dropout))
        self.nn.append(nn.Linear(widths[-2], widths[-1]))

    @property
    def n_params(self) -> int:
        return sum(p.numel() for p in self.parameters())

    def linears(self) -> list[nn.Linear]:
        return [m for m in self.nn if isinstance(m, nn.Linear)]

    def to(self, device: Any) -> "ConditionalNestedNeuralNetwork":  # type: ignore[override]
        super().to(device)
        self.device = device
        return self

    def forward(self, y: torch.Tensor, h: torch.Tensor):  # pragma: no cover - not an entry point
        raise NotImplementedError("the conditioner is fused into the coupling kernels; call the coupling layer "
                                  "or the flow instead")


class _SingleLayerMixin:
    """Lets one layer be called on its own (the reference's unit test does, tests/test_cnf.py:18-32)."""
    _packed: PackedFlow | None = None

    def _pack_self(self, size: int, n_conditions: int, nested: Sequence[int], two_way: bool) -> PackedFlow:
        dev = _as_device(self.device)
        if self._packed is None or self._packed.device != dev:
            object.__setattr__(self, "_packed", PackedFlow([self], size, n_conditions, nested, two_way, dev))
        return self._packed


class ConditionalAffineCouplingLayer(ConditionalInvertibleLayer, _SingleLayerMixin):
    """Reference cnf.py:110-213."""

    def __init__(self, input_size: int, nested_sizes: list[int], n_conditions: int, layer: str = "Linear",
                 layer_kwargs: dict[str, Any] | None = None, activation: str = "GELU",
                 activation_kwargs: dict[str, Any] | None = None, dropout: float = 0.0, device: Any = "cpu",
                 two_way: bool = False) -> None:
        super().__init__()
        self.input_size = input_size
        self.nested_sizes = list(nested_sizes)
        self.n_conditions = n_conditions
        self.log_det_J: torch.Tensor | float = torch.zeros(1)
        self.device = device
        self.two_way = two_way
        d_a, d_b = int(math.ceil(input_size / 2)), int(math.floor(input_size / 2))   # cnf.py:136
        self.nn_a = ConditionalNestedNeuralNetwork([d_a] + self.nested_sizes + [d_b], n_conditions, 2, layer,
                                                   layer_kwargs, activation, activation_kwargs, dropout, device)
        if two_way:
            self.nn_b = ConditionalNestedNeuralNetwork([d_b] + self.nested_sizes + [d_a], n_conditions, 2, layer,
                                                       layer_kwargs, activation, activation_kwargs, dropout, device)

    def to(self, device: Any) -> "ConditionalAffineCouplingLayer":  # type: ignore[override]
        super().to(device)
        self.device = device
        return self

    def _run(self, inverse: bool, v: torch.Tensor, h: torch.Tensor, want_ld: bool):
        if v.dim() == 1:
            v = v.unsqueeze(0)       # cnf.py:167-172
        if h.dim() == 1:
            h = h.unsqueeze(0)
        pk = self._pack_self(self.input_size, self.n_conditions, self.nested_sizes, self.two_way)
        if h.shape[0] != v.shape[0]:
            raise ValueError(f"got {v.shape[0]} rows but {h.shape[0]} condition rows")
        return pk.run(inverse, v, pk.project(h), want_logdet=want_ld)

    def forward(self, y: torch.Tensor, x: torch.Tensor, log_det_J: bool = False) -> torch.Tensor:
        z, ld = self._run(False, y, x, log_det_J)
        if log_det_J:
            self.log_det_J = ld
        return z

    def inverse(self, z: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return self._run(True, z, y, False)[0]


class OrthonormalTransformation(ConditionalInvertibleLayer, _SingleLayerMixin):
    """Reference cnf.py:312-339: a fixed random orthonormal matrix from a QR decomposition."""

    def __init__(self, input_size: int, random_state: int | None = None) -> None:
        super().__init__()
        self.input_size = input_size
        self.log_det_J: float = 0
        self.device: Any = "cpu"
        if random_state is not None:
            # Reference behaviour, kept: reseeds the GLOBAL generator before every construction,
            # so all mixing matrices of a model are identical when random_state is set (cnf.py:319-320).
            torch.manual_seed(random_state)
        q = torch.linalg.qr(torch.randn(input_size, input_size))[0]
        self.orthonormal_matrix = nn.Parameter(q, requires_grad=False)

    def to(self, device: Any) -> "OrthonormalTransformation":  # type: ignore[override]
        super().to(device)
        self.device = device
        return self

    def forward(self, y: torch.Tensor, x: torch.Tensor | None = None, log_det_J: bool = False) -> torch.Tensor:
        pk = self._pack_self(self.input_size, 1, [16], False)
        return pk.run(False, y, torch.zeros(1, 1, device=pk.device))[0]

    def inverse(self, z: torch.Tensor, x: torch.Tensor | None = None) -> torch.Tensor:
        pk = self._pack_self(self.input_size, 1, [16], False)
        return pk.run(True, z, torch.zeros(1, 1, device=pk.device))[0]


class ActNorm(InvertibleLayer, _SingleLayerMixin):
    """Reference cnf.py:342-354: per-dimension scale and bias, no data-dependent init."""

    def __init__(self, size: int) -> None:
        super().__init__()
        self.size = size
        self.device: Any = "cpu"
        self.scale = nn.Parameter(torch.ones(size))
        self.bias = nn.Parameter(torch.zeros(size))

    def to(self, device: Any) -> "ActNorm":  # type: ignore[override]
        super().to(device)
        self.device = device
        return self

    def forward(self, x: torch.Tensor, log_det_J: bool = False) -> torch.Tensor:
        pk = self._pack_self(self.size, 1, [16], False)
        z, ld = pk.run(False, x, torch.zeros(1, 1, device=pk.device), want_logdet=True)
        self.log_det_J = ld[0] if ld.numel() else torch.zeros((), device=pk.device)   # set regardless of the flag, cnf.py:350
        return z

    def inverse(self, z: torch.Tensor) -> torch.Tensor:
        pk = self._pack_self(self.size, 1, [16], False)
        return pk.run(True, z, torch.zeros(1, 1, device=pk.device))[0]


# ------------------------------------------------------------------------------------------
# the flow
# ------------------------------------------------------------------------------------------
class CondRealNVP_v2(ConditionalInvertibleLayer):
    """Reference cnf.py:357-588 -- same constructor, methods and state_dict; B200 kernels beneath.

    Extra keyword-only knobs (defaults keep the reference's call signatures valid):

    ``precision``: arithmetic of the conditioner GEMMs -- ``"fp32"`` (FMA kernels), ``"bf16x3"``
    (tcgen05, 3-term bf16 split, fp32-class accuracy), ``"bf16"`` (tcgen05, one bf16 pass, stated
    tolerance) or ``"auto"`` (bf16x3 where the conditioner is wide enough, else fp32).

    ``sample_rng``: ``"device"`` (default) draws z on the GPU and evaluates all rows of an
    instance chunk in one launch; ``"reference"`` replays the reference's loops and draws z from
    the CPU generator exactly as cnf.py:566/:578/:584 do, so a seeded run reproduces the
    reference's samples to fp32 tolerance.
    """

    #: conditioning instances whose projection P is kept resident per launch (bounds memory)
    max_proj_bytes = 2 << 30

    def __init__(self, size: int, nested_sizes: list[int], n_blocks: int, n_conditions: int,
                 feature_networks: list[FeatureNetwork | nn.Module | None] | None = None, dropout: float = 0.0,
                 act_norm: bool = False, two_way: bool = False, layer: str = "Linear",
                 layer_kwargs: dict[str, Any] | None = None, activation: str = "GELU",
                 activation_kwargs: dict[str, Any] | None = None, device: Any = "cpu",
                 random_state: int | None = None, parameter_index_mapping: ParameterIndexMapping | None = None,
                 hybrid: bool = False, *, sample_rng: str = "device", precision: str = "auto") -> None:
        super().__init__()
        if n_conditions <= 0:
            # the reference accepts n_conditions == 0 in the constructor but every forward then
            # raises ValueError (cnf.py:378-379, :480-485); refuse at construction instead
            raise NotImplementedError("n_conditions must be > 0 (the reference's unconditional path is broken)")
        _require_linear_gelu(layer, layer_kwargs, activation, activation_kwargs)
        if sample_rng not in ("device", "reference"):
            raise ValueError(f"sample_rng must be 'device' or 'reference', got {sample_rng!r}")
        self.feature_network_stack = FeatureNetworkStack(feature_networks)
        self.size = size
        self.nested_sizes = list(nested_sizes)
        self.n_blocks = n_blocks
        self.n_conditions = n_conditions
        self.device = device
        self.dropout = dropout
        self.act_norm = act_norm
        self.two_way = two_way
        self.parameter_index_mapping = parameter_index_mapping
        self.hybrid = hybrid
        self.sample_rng = sample_rng
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
        self.precision = precision
        self.log_det_J: torch.Tensor = torch.zeros(1)
        if hybrid:
            self.prediction_head = nn.Linear(n_conditions, size)     # cnf.py:391-392

        def coupling() -> ConditionalAffineCouplingLayer:
            return ConditionalAffineCouplingLayer(size, self.nested_sizes, n_conditions, layer=layer,
                                                  layer_kwargs=layer_kwargs, activation=activation,
                                                  activation_kwargs=activation_kwargs, dropout=dropout,
                                                  two_way=two_way, device=device)

        # [ActNorm?, coupling, orthonormal] x (n_blocks - 1) + coupling   (cnf.py:395-423)
        self.layers = nn.ModuleList()
        for _ in range(n_blocks - 1):
            if act_norm:
                self.layers.append(ActNorm(size))
            self.layers.append(coupling())
            self.layers.append(OrthonormalTransformation(size, random_state=random_state))
        self.layers.append(coupling())
        self._packed: PackedFlow | None = None

    # -- construction helpers -------------------------------------------------------------
    def verify(self) -> None:
        """Feature-network sizes must chain and end at n_conditions (cnf.py:425-440)."""
        current = None
        for fn in self.feature_network_stack.feature_networks:
            if isinstance(fn, FeatureNetwork):
                if current is not None:
                    assert current == fn.input_size, (
                        "The output dimension of the feature network must match the input dimension of the time "
                        f"series network. Have {current} but need {fn.input_size} for next layer.")
                current = fn.output_size
        if current is not None:
            assert current == self.n_conditions, (
                "The output dimension of the time series network must match the number of conditions. "
                f"Have {current} but need {self.n_conditions}.")

    @classmethod
    def from_config(cls, config: dict[str, Any], **extra: Any) -> "CondRealNVP_v2":
        """Build from a run config (cnf.py:442-456); accepts plain dicts and Dynaconf-like objects."""
        nets = [FeatureNetworkFactory.get_feature_network(fc["type"], dict(fc.get("kwargs", {}) or {}))
                for fc in config["feature_networks"]]
        kwargs = dict(config["model"]["kwargs"])
        kwargs.update(extra)
        cnf = cls(feature_networks=nets,
                  parameter_index_mapping=ParameterIndexMapping(list(config["global"]["parameter_selection"])),
                  **kwargs)
        cnf.verify()
        return cnf

    def to(self, device: Any) -> "CondRealNVP_v2":  # type: ignore[override]
        super().to(device)
        self.device = device
        for layer in self.layers:
            layer.device = device
        return self

    @property
    def n_params(self) -> int:
        return sum(p.numel() for p in self.parameters())

    # -- plumbing ----------------------------------------------------------------------
    def _flow(self) -> PackedFlow:
        dev = _as_device(self.device)
        if self._packed is None or self._packed.device != dev:
            self._packed = PackedFlow(list(self.layers), self.size, self.n_conditions, self.nested_sizes,
                                      self.two_way, dev, self.precision)
        return self._packed

    def _check_mode(self, what: str) -> None:
        if self.training and self.dropout > 0.0:
            raise NotImplementedError(
                f"{what} in training mode would run the conditioner with dropout (p={self.dropout}, cnf.py:82-83); "
                "only forward() has a training path -- call .eval() before inverse() / sample()")

    def features(self, *conditions: torch.Tensor) -> torch.Tensor:
        """Condition features h = feature_network_stack(*conditions) on the model's device.

        On tensor-core handles the FullyConnected / LSTM feature networks follow the stack's arithmetic mode whenever the
        model is in eval mode (bcnf_b200/feature_tc.py; no autograd history -- the eval-mode stack has none either);
        in training mode, and for the Transformer, the plain PyTorch module runs.
        """
        dev = _as_device(self.device)
        if dev.type == "cuda" and not self.training:
            flow = self._flow()
            self._tc_passes(flow)
        return self.feature_network_stack(*[c.to(dev) for c in conditions])

    def _tc_passes(self, flow: PackedFlow) -> int:
        """Arithmetic mode of the handle (3 = bf16x3, 1 = bf16, 0 = no tensor-core path), handed on to the encoders."""
        passes = {"bf16x3": 3, "bf16": 1}.get(flow.precision, 0) if flow.kernel == "tcgen05" else 0
        for fn in self.feature_network_stack.feature_networks:
            if hasattr(fn, "tc_passes"):
                fn.tc_passes = passes
        return passes

    def _projection(self, *conditions: torch.Tensor) -> torch.Tensor:
        """Raw conditions -> P (n_inst, proj_width), the hoisted condition terms of every conditioner's first Linear.

        Eval mode on a tensor-core handle: the last feature network's affine output layer and the projection are ONE GEMM
        (feature_tc.fused_projection; h is never materialised); otherwise features() followed by bcnf_cond_project."""
        flow = self._flow()
        dev = _as_device(self.device)
        if not self.training and not self.feature_network_stack.training:
            passes = self._tc_passes(flow)
            if passes:
                from . import feature_tc
                P = feature_tc.fused_projection(self, flow, tuple(c.to(dev) for c in conditions), passes)
                if P is not None:
                    return P
        with torch.no_grad():
            return flow.project(self.features(*conditions))      # tensor-core handles: img_pack + GEMM; fp32: one SGEMM

    # -- reference API ------------------------------------------------------------------
    def forward(self, y: torch.Tensor, *conditions: torch.Tensor, log_det_J: bool = False,
                return_features: bool = False, deterministic_features: bool = False):
        """cnf.py:467-493.  ``self.log_det_J`` is set as a side effect when ``log_det_J=True``.

        In training mode (``model.train()``) the call goes through the differentiable path of
        ``bcnf_b200/train.py`` (conditioner dropout active, autograd history on z and ``log_det_J``); in
        eval mode through the fused inference kernels, without autograd history.
        """
        if not self.training and not return_features and not deterministic_features:
            # eval mode, h not asked for: features and projection fused (the log-prob / NLL evaluation path)
            P = self._projection(*conditions)
            if P.shape[0] != y.shape[0]:
                raise ValueError(f"got {y.shape[0]} rows but {P.shape[0]} condition rows")
            flow = self._flow()
            with torch.no_grad():
                z, ld = flow.run(False, y, P, want_logdet=log_det_J)
            if log_det_J:
                self.log_det_J = ld
            return z
        if deterministic_features:
            self.feature_network_stack.eval()
            condition = self.features(*conditions).detach()
        else:
            condition = self.features(*conditions)
        if condition.shape[0] != y.shape[0]:
            raise ValueError(f"got {y.shape[0]} rows but {condition.shape[0]} condition rows")
        if self.training:
            from .train import stack_forward_train
            z, ld = stack_forward_train(self, y, condition)
            if log_det_J:
                self.log_det_J = ld
            return (z, condition) if return_features else z
        flow = self._flow()
        with torch.no_grad():
            z, ld = flow.run(False, y, flow.project(condition), want_logdet=log_det_J)
        if log_det_J:
            self.log_det_J = ld
        if return_features:
            return z, condition
        return z

    def inverse(self, z: torch.Tensor, *conditions: torch.Tensor) -> torch.Tensor:
        """cnf.py:495-508."""
        self._check_mode("inverse")
        P = self._projection(*conditions)
        if P.shape[0] != z.shape[0]:
            raise ValueError(f"got {z.shape[0]} rows but {P.shape[0]} condition rows")
        flow = self._flow()
        with torch.no_grad():
            return flow.run(True, z, P)[0]

    def log_prob(self, y: torch.Tensor, *conditions: torch.Tensor, reference_scale: bool = False) -> torch.Tensor:
        """New convenience (the reference has no log_prob).

        ``log p(y | c) = -0.5 sum z^2 + log|det J| - D/2 log(2 pi)``.  ``reference_scale=True`` drops
        the constant, i.e. returns ``-inn_nll_loss(z, log_det_J, reduction='none')`` (utils.py:49-53).
        """
        z = self.forward(y, *conditions, log_det_J=True)
        lp = -0.5 * (z * z).sum(dim=1) + self.log_det_J
        if not reference_scale:
            lp = lp - 0.5 * self.size * math.log(2.0 * math.pi)
        return lp

    def sample(self, n_samples: int, *conditions: torch.Tensor, sigma: float = 1, outer: bool = False,
               batch_size: int = 100, sample_batch_size: int | None = None, output_device: Any = "cpu",
               verbose: bool = False) -> torch.Tensor:
        """cnf.py:510-538: (n_samples, N, D) for ``outer=True``."""
        self._check_mode("sample")
        if self.sample_rng == "reference" or not outer or not all(c.ndim > 1 for c in conditions):
            return self._sample_reference_loops(n_samples, *conditions, sigma=sigma, outer=outer,
                                                batch_size=batch_size, sample_batch_size=sample_batch_size,
                                                output_device=output_device)
        return self._sample_device(n_samples, *conditions, sigma=sigma, output_device=output_device)

    # reference-order sampling: same loops, same CPU generator draws (cnf.py:510-538)
    def _sample_reference_loops(self, n_samples: int, *conditions: torch.Tensor, sigma: float, outer: bool,
                                batch_size: int, sample_batch_size: int | None, output_device: Any) -> torch.Tensor:
        if sample_batch_size is None:
            sample_batch_size = batch_size
        m_batch_sizes = [sample_batch_size] * (n_samples // sample_batch_size) + [n_samples % sample_batch_size]
        rows: list[torch.Tensor] = []
        with torch.no_grad():
            for b in range(0, len(conditions[0]), batch_size):
                batch_conditions = [c[b: b + batch_size] for c in conditions]
                parts = [self._sample(m, *batch_conditions, outer=outer, sigma=sigma).to(output_device)
                         for m in m_batch_sizes if m != 0]
                rows.append(torch.cat(parts, dim=0))
        return torch.cat(rows, dim=1)

    def _sample(self, n_samples: int, *conditions: torch.Tensor, sigma: float = 1, outer: bool = False,
                z: torch.Tensor | None = None) -> torch.Tensor:
        """cnf.py:540-588, the three condition modes.  z is drawn on the CPU generator like the
        reference unless injected; features are computed once per instance, not per row."""
        dev = _as_device(self.device)
        flow = self._flow()
        with torch.no_grad():
            if all(c.ndim == 1 for c in conditions):                      # cnf.py:564-570
                h = self.features(*[c.unsqueeze(0) for c in conditions])
                if z is None:
                    z = sigma * torch.randn(n_samples, self.size)
                x, _ = flow.run(True, z.to(dev), flow.project(h), inst_period=1)
                return x.view(n_samples, self.size)
            if all(c.ndim > 1 for c in conditions):
                if outer:                                                 # cnf.py:572-582
                    if not len(set(c.shape[0] for c in conditions)) == 1:
                        raise ValueError("All conditions must have the same number of samples (dim = 0). "
                                         f"Got {[c.shape for c in conditions]}.")
                    n_inst = conditions[0].shape[0]
                    h = self.features(*conditions)
                    if z is None:
                        z = sigma * torch.randn(n_samples * n_inst, self.size)
                    x, _ = flow.run(True, z.to(dev), flow.project(h), inst_period=n_inst)
                    return x.view(n_samples, n_inst, self.size)
                h = self.features(*conditions)                            # cnf.py:583-586
                if z is None:
                    z = sigma * torch.randn(n_samples, self.size)
                if h.shape[0] != n_samples:
                    raise ValueError(f"outer=False needs one condition row per sample: got {h.shape[0]} for {n_samples}")
                x, _ = flow.run(True, z.to(dev), flow.project(h))
                return x.view(n_samples, self.size)
        raise ValueError(f"Conditions have invalid shape: {[c.shape for c in conditions]}")   # cnf.py:588

    # fast path: features and projection once per instance, all rows of a chunk in one launch
    def _sample_device(self, n_samples: int, *conditions: torch.Tensor, sigma: float, output_device: Any,
                       generator: torch.Generator | None = None) -> torch.Tensor:
        if not len(set(c.shape[0] for c in conditions)) == 1:
            raise ValueError("All conditions must have the same number of samples (dim = 0). "
                             f"Got {[c.shape for c in conditions]}.")
        dev = _as_device(self.device)
        flow = self._flow()
        n_inst = conditions[0].shape[0]
        out_dev = _as_device(output_device)
        out = torch.empty((n_samples, n_inst, self.size), dtype=torch.float32, device=out_dev,
                          pin_memory=(out_dev.type == "cpu"))
        if n_inst == 0 or n_samples == 0:
            return out
        per_inst = 4 * max(flow.proj_width, 1)
        chunk = max(1, min(n_inst, self.max_proj_bytes // per_inst, max(1, (1 << 27) // max(n_samples, 1))))
        with torch.no_grad():
            for b in range(0, n_inst, chunk):
                cs = [c[b: b + chunk] for c in conditions]
                nb = cs[0].shape[0]
                P = self._projection(*cs)
                # z = sigma * N(0, 1) is drawn inside the kernel (Philox keyed by one seed per chunk, taken from the
                # generator like a torch.randn call would advance it): no z tensor, no torch.randn launch
                seed = int(torch.randint(0, 2 ** 62, (1,), generator=generator, device=generator.device).item()) \
                    if generator is not None else int(torch.randint(0, 2 ** 62, (1,)).item())
                x = flow.sample(n_samples * nb, P, seed=seed, sigma=sigma, inst_period=nb)
                out[:, b: b + nb].copy_(x.view(n_samples, nb, self.size), non_blocking=True)
        if out_dev.type == "cpu":
            torch.cuda.current_stream(dev).synchronize()
        return out
