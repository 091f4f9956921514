"""FullyConnectedFeatureNetwork on the tensor cores (inference): SURVEY.md section 8f-1, reference
src/bcnf/models/feature_network.py:114-145 (Linear -> GELU [-> Dropout] ... -> Linear).

For log-prob / NLL evaluation every row has its own condition, so the feature MLP runs on as many rows as the
stack (FC_large: 1.03 M MACs per instance) and an fp32 SIMT GEMM chain costs a sixth of the step.  Here the MLP is
a chain of CTA-pair GEMMs on operand images (csrc/gemm_img2.cuh): x -> image; every hidden layer
gelu(a W^T + b) -> image; last layer -> fp32 h.  Same arithmetic modes as the stack (3-pass bf16 split = fp32-class,
or single-pass bf16).  Eval mode, no autograd; parameter images are cached per (tensor, version).
"""
from __future__ import annotations

import ctypes as C
from typing import Any

import torch
from torch import nn

from . import _cabi
from .train import _Img, _img, _pack_images, _stream

MIN_ROWS = 2048     # below this the nine launches cost more than the PyTorch chain


def supported(net: Any) -> bool:
    mods = list(net.nn)
    if not mods or not isinstance(mods[-1], nn.Linear):
        return False
    if any(p.dtype != torch.float32 for p in net.parameters()):
        return False
    for m in mods:
        if isinstance(m, nn.GELU):
            if m.approximate != "none":
                return False
        elif not isinstance(m, (nn.Linear, nn.Dropout)):
            return False
    # every Linear but the last must be followed by GELU
    lin = [i for i, m in enumerate(mods) if isinstance(m, nn.Linear)]
    return all(isinstance(mods[i + 1], nn.GELU) for i in lin[:-1])


def _weight_image(net: Any, lin: nn.Linear) -> _Img:
    cache = net.__dict__.setdefault("_tc_images", {})
    w = lin.weight
    key = (id(lin), w.data_ptr(), w._version, tuple(w.shape))
    hit = cache.get(id(lin))
    if hit is not None and hit[0] == key:
        return hit[1]
    im = _Img(w.device, w.shape[0], w.shape[1], align=256)
    wc = w.detach().contiguous()
    _pack_images([(wc, 0, wc.stride(0), 1, w.shape[0], w.shape[1], im)], w.device)
    cache[id(lin)] = (key, im)
    return im


def forward(net: Any, x: torch.Tensor, passes: int) -> torch.Tensor:
    """x (rows, features) on a CUDA device -> h (rows, output_size), fp32."""
    dev = x.device
    x = x.reshape(x.size(0), -1).contiguous().float()
    rows = x.shape[0]
    lib = _cabi.lib()
    linears = [m for m in net.nn if isinstance(m, nn.Linear)]
    a = _img(dev, ("fc_in", id(net)), rows, x.shape[1], align=256)
    _pack_images([(x, 0, x.stride(0), 1, rows, x.shape[1], a)], dev)
    for li, lin in enumerate(linears):
        wimg = _weight_image(net, lin)
        n_out, n_in = lin.weight.shape
        bias = lin.bias.detach() if lin.bias is not None else torch.zeros(n_out, device=dev)
        if li < len(linears) - 1:
            c = _img(dev, ("fc_act", id(net), li & 1), rows, n_out, align=256)
            _cabi.check(lib.bcnf_gemm_img_gelu(a.ptr, a.plane, a.rpad, wimg.ptr, wimg.plane, wimg.rpad, bias.data_ptr(), c.ptr,
                                               c.plane, c.rpad, rows, n_out, n_in, passes, dev.index or 0, _stream(dev)),
                        "bcnf_gemm_img_gelu")
            a = c
        else:
            h = torch.empty(rows, n_out, device=dev)
            _cabi.check(lib.bcnf_gemm_img(a.ptr, a.plane, a.rpad, wimg.ptr, wimg.plane, wimg.rpad, h.data_ptr(), n_out,
                                          bias.data_ptr(), rows, n_out, n_in, passes, dev.index or 0, _stream(dev)),
                        "bcnf_gemm_img")
            return h
    raise AssertionError("unreachable")
