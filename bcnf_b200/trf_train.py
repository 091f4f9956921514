"""Transformer condition encoder inside a Trainer step (reference src/bcnf/models/feature_network.py:183-307;
SURVEY.md section 8f-1 / 8a12): forward on the package's own kernels, backward written out by hand.

At batch 256 the PyTorch module is a dependency chain of ~250 launch-bound kernels forward (a Linear alone is three:
split-K SGEMM, reduction, bias epilogue) and ~300 backward: 2.6 + 2.3 ms of an 11 ms step.  Here the forward is the
inference path of feature_tc.transformer_token0 -- eight launches per block: q | k | v GEMM, attention, fc_out GEMM,
dropout + add + LayerNorm, FFN GEMM, GELU, FFN GEMM, dropout + add + LayerNorm -- with the tensors the backward needs
written out by the same kernels, and the backward is eleven launches per block on the chain (LayerNorm and GELU through
ATen's backward kernels, the attention core through bcnf_trf_attention_bwd, data gradients as cuBLAS GEMMs); every
parameter gradient is computed on the Trainer's side streams (feature_network.OffChain) and written straight into .grad.

Dropout (nn.Dropout(p) on the embedding, on both sublayer outputs of every block and on the final state) uses
multiplier tensors drawn once per step from torch's generator (graph-safe Philox): the same tensors scale the forward
values inside the kernels and the gradients in the backward.
"""
from __future__ import annotations

import os
from typing import Any

import torch

from . import _cabi
from .feature_network import _wgrad
from .feature_tc import _img, _Img, _pack_images, _stream, transformer_supported

ENABLED = os.environ.get("BCNF_TRAIN_TRF_KERNELS", "1") != "0"
_aten = torch.ops.aten


def usable(net: Any, x: torch.Tensor) -> bool:
    return (ENABLED and x.is_cuda and x.ndim == 3 and x.shape[1] <= 32 and x.dtype == torch.float32
            and transformer_supported(net) and all(blk.ffn[0].out_features % 8 == 0 for blk in net.layers)
            and all(p.requires_grad for p in net.parameters()))


def _mask(shape: tuple, p: float, dev: torch.device) -> torch.Tensor | None:
    """nn.Dropout(p) as a tensor of multipliers: 0 with probability p, 1 / (1 - p) otherwise."""
    if p <= 0.0:
        return None
    return torch.empty(shape, device=dev).bernoulli_(1.0 - p).mul_(1.0 / (1.0 - p))


def _block_masks(net: Any, rows: int, E: int, dev: torch.device) -> list:
    """(m1, m2) per block.  Blocks that share a dropout probability (all of them, as the constructor builds them) draw
    their multipliers with ONE generator call: two launches per step instead of four per block."""
    ps = [float(blk.dropout.p) for blk in net.layers]
    if len(set(ps)) == 1:
        m = _mask((2 * len(ps), rows, E), ps[0], dev)
        return [(None, None) if m is None else (m[2 * l], m[2 * l + 1]) for l in range(len(ps))]
    return [(_mask((rows, E), p, dev), _mask((rows, E), p, dev)) for p in ps]


def _weight_images(net: Any, wqkv_all: torch.Tensor) -> list[dict]:
    """Operand images of the blocks' matrices, repacked every step (the optimizer has just changed them) by one launch."""
    dev = wqkv_all.device
    E = net.trf_size
    imgs = net.__dict__.get("_trf_train_imgs")
    if imgs is None or imgs[0]["wqkv"].buf.device != dev:
        imgs = [{"wqkv": _Img(dev, 3 * E, E, align=256), "wo": _Img(dev, E, E, align=256),
                 "w1": _Img(dev, blk.ffn[0].out_features, E, align=256),
                 "w2": _Img(dev, E, blk.ffn[0].out_features, align=256)} for blk in net.layers]
        net.__dict__["_trf_train_imgs"] = imgs
    descs = []
    for l, (blk, im) in enumerate(zip(net.layers, imgs)):
        wo, w1, w2 = blk.attention.fc_out.weight.detach(), blk.ffn[0].weight.detach(), blk.ffn[2].weight.detach()
        descs.append((wqkv_all, l * 3 * E * E, E, 1, 3 * E, E, im["wqkv"]))
        descs.append((wo, 0, wo.stride(0), 1, wo.shape[0], wo.shape[1], im["wo"]))
        descs.append((w1, 0, w1.stride(0), 1, w1.shape[0], w1.shape[1], im["w1"]))
        descs.append((w2, 0, w2.stride(0), 1, w2.shape[0], w2.shape[1], im["w2"]))
    for i in range(0, len(descs), 64):
        _pack_images(descs[i: i + 64], dev)
    return imgs


class _TransformerTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net: Any, oc: Any, tokens: torch.Tensor, *params: torch.Tensor):
        dev = tokens.device
        tokens = tokens.contiguous()
        B, T, F = tokens.shape
        E, L = net.trf_size, len(net.layers)
        rows = B * T
        lib, di, st = _cabi.lib(), dev.index or 0, _stream(dev)
        new = lambda *shape: torch.empty(*shape, device=dev)
        with torch.no_grad():
            wqkv_all = torch.cat([w for blk in net.layers for w in (blk.attention.q_linear.weight, blk.attention.k_linear.weight,
                                                                  blk.attention.v_linear.weight)], dim=0)
            bqkv_all = torch.cat([b for blk in net.layers for b in (blk.attention.q_linear.bias, blk.attention.k_linear.bias,
                                                                  blk.attention.v_linear.bias)], dim=0)
            imgs = _weight_images(net, wqkv_all)
            p_io = float(net.dropout.p)
            m_in, m_out = _mask((rows, E), p_io, dev), _mask((B, E), p_io, dev)
            block_masks = _block_masks(net, rows, E, dev)
            pos = net._positional(T, dev).contiguous() if net.add_positional_embeddings else None
            tag = ("trf_train", id(net), rows)
            x_img = _img(dev, (tag, "x"), rows, E, align=256)
            c_img = _img(dev, (tag, "ctx"), rows, E, align=256)
            fw, fb = net.features.weight, net.features.bias
            x = new(rows, E)
            _cabi.check(lib.bcnf_trf_embed(tokens.data_ptr(), fw.data_ptr(), fb.data_ptr(), pos.data_ptr() if pos is not None else None,
                                           m_in.data_ptr() if m_in is not None else None, rows, T, F, E, x.data_ptr(),
                                           x_img.ptr, x_img.plane, x_img.rpad, di, st), "bcnf_trf_embed")
            saved = []
            for l, (blk, im) in enumerate(zip(net.layers, imgs)):
                heads, ff = blk.attention.n_heads, blk.ffn[0].out_features
                m1, m2 = block_masks[l]
                f_img = _img(dev, (tag, "ffn", ff), rows, ff, align=256)
                n1, n2 = blk.norm1, blk.norm2
                qkv, cx, o = new(rows, 3 * E), new(rows, E), new(rows, E)
                _cabi.check(lib.bcnf_gemm_img(x_img.ptr, x_img.plane, x_img.rpad, im["wqkv"].ptr, im["wqkv"].plane, im["wqkv"].rpad,
                                              qkv.data_ptr(), 3 * E, bqkv_all[l * 3 * E:].data_ptr(), rows, 3 * E, E, 3, di, st),
                            "bcnf_gemm_img")
                _cabi.check(lib.bcnf_trf_attention(qkv.data_ptr(), B, T, E, heads, cx.data_ptr(), c_img.ptr, c_img.plane,
                                                   c_img.rpad, di, st), "bcnf_trf_attention")
                _cabi.check(lib.bcnf_gemm_img(c_img.ptr, c_img.plane, c_img.rpad, im["wo"].ptr, im["wo"].plane, im["wo"].rpad,
                                              o.data_ptr(), E, blk.attention.fc_out.bias.data_ptr(), rows, E, E, 3, di, st),
                            "bcnf_gemm_img")
                s1, mean1, rstd1, x1 = new(rows, E), new(rows), new(rows), new(rows, E)
                _cabi.check(lib.bcnf_trf_add_layernorm(x.data_ptr(), o.data_ptr(), m1.data_ptr() if m1 is not None else None,
                                                       n1.weight.data_ptr(), n1.bias.data_ptr(), n1.eps, rows, E, s1.data_ptr(),
                                                       mean1.data_ptr(), rstd1.data_ptr(), x1.data_ptr(), x_img.ptr, x_img.plane,
                                                       x_img.rpad, di, st), "bcnf_trf_add_layernorm")
                u, a, f = new(rows, ff), new(rows, ff), new(rows, E)
                _cabi.check(lib.bcnf_gemm_img(x_img.ptr, x_img.plane, x_img.rpad, im["w1"].ptr, im["w1"].plane, im["w1"].rpad,
                                              u.data_ptr(), ff, blk.ffn[0].bias.data_ptr(), rows, ff, E, 3, di, st), "bcnf_gemm_img")
                _cabi.check(lib.bcnf_trf_gelu(u.data_ptr(), rows, ff, a.data_ptr(), f_img.ptr, f_img.plane, f_img.rpad, di, st),
                            "bcnf_trf_gelu")
                _cabi.check(lib.bcnf_gemm_img(f_img.ptr, f_img.plane, f_img.rpad, im["w2"].ptr, im["w2"].plane, im["w2"].rpad,
                                              f.data_ptr(), E, blk.ffn[2].bias.data_ptr(), rows, E, ff, 3, di, st), "bcnf_gemm_img")
                s2, mean2, rstd2, x2 = new(rows, E), new(rows), new(rows), new(rows, E)
                _cabi.check(lib.bcnf_trf_add_layernorm(x1.data_ptr(), f.data_ptr(), m2.data_ptr() if m2 is not None else None,
                                                       n2.weight.data_ptr(), n2.bias.data_ptr(), n2.eps, rows, E, s2.data_ptr(),
                                                       mean2.data_ptr(), rstd2.data_ptr(), x2.data_ptr(), x_img.ptr, x_img.plane,
                                                       x_img.rpad, di, st), "bcnf_trf_add_layernorm")
                saved.append((x, qkv, cx, s1, mean1, rstd1, m1, x1, u, a, s2, mean2, rstd2, m2))
                x = x2
            t0 = x.view(B, T, E)[:, 0, :]
            t0 = t0 * m_out if m_out is not None else t0.contiguous()
            h = torch.addmm(net.output.bias, t0, net.output.weight.t())
        ctx.net, ctx.oc, ctx.saved = net, oc, saved
        ctx.misc = (tokens, m_in, m_out, t0, wqkv_all, B, T, E)
        ctx.param_versions = [p._version for p in params]
        return h

    @staticmethod
    def backward(ctx, dh: torch.Tensor):
        net, oc = ctx.net, ctx.oc
        tokens, m_in, m_out, t0, wqkv_all, B, T, E = ctx.misc
        if [p._version for p in net.parameters()] != ctx.param_versions:
            raise RuntimeError("bcnf_b200: a parameter of the Transformer encoder was modified in place between the forward "
                               "and the backward pass of a training step")
        dev = dh.device
        rows = B * T
        lib, di, st = _cabi.lib(), dev.index or 0, _stream(dev)

        def ln_bwd(g, s, mean, rstd, ln):
            """dx of nn.LayerNorm on the chain; d gamma / d beta (column reductions only the optimizer reads) off it."""
            mean2, rstd2 = mean.view(rows, 1), rstd.view(rows, 1)
            dx = _aten.native_layer_norm_backward(g, s, [E], mean2, rstd2, ln.weight, ln.bias, [True, False, False])[0]

            def params_grad():
                dwb = torch.zeros(2, E, device=dev)
                _cabi.check(lib.bcnf_trf_ln_param_grad(g.data_ptr(), s.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rows, E,
                                                       dwb[0].data_ptr(), dwb[1].data_ptr(), di, _stream(dev)),
                            "bcnf_trf_ln_param_grad")
                oc.accumulate(ln.weight, dwb[0])
                oc.accumulate(ln.bias, dwb[1])
            oc.run(params_grad, g, s, mean, rstd)
            return dx

        with torch.no_grad():
            dh = dh.contiguous()
            out_lin = net.output
            dt0 = dh.mm(out_lin.weight)
            oc.defer(t0, dh, out_lin.weight, out_lin.bias)
            if m_out is not None:
                dt0 = dt0 * m_out
            dx = torch.zeros(B, T, E, device=dev)
            dx[:, 0, :] = dt0
            dx = dx.view(rows, E)
            for l in range(len(net.layers) - 1, -1, -1):
                blk = net.layers[l]
                x_in, qkv, cx, s1, mean1, rstd1, m1, x1, u, a, s2, mean2, rstd2, m2 = ctx.saved[l]
                att, lin1, lin2 = blk.attention, blk.ffn[0], blk.ffn[2]
                ds2 = ln_bwd(dx, s2, mean2, rstd2, blk.norm2)
                df = ds2 * m2 if m2 is not None else ds2
                da = df.mm(lin2.weight)
                oc.defer(a, df, lin2.weight, lin2.bias)
                du = _aten.gelu_backward(da, u, approximate="none")
                dx1 = torch.addmm(ds2, du, lin1.weight)
                oc.defer(x1, du, lin1.weight, lin1.bias)
                ds1 = ln_bwd(dx1, s1, mean1, rstd1, blk.norm1)
                do = ds1 * m1 if m1 is not None else ds1
                dctx = do.mm(att.fc_out.weight)
                oc.defer(cx, do, att.fc_out.weight, att.fc_out.bias)
                dqkv = torch.empty(rows, 3 * E, device=dev)
                _cabi.check(lib.bcnf_trf_attention_bwd(qkv.data_ptr(), dctx.data_ptr(), B, T, E, att.n_heads, dqkv.data_ptr(), di, st),
                            "bcnf_trf_attention_bwd")
                dx = torch.addmm(ds1, dqkv, wqkv_all[l * 3 * E: (l + 1) * 3 * E])

                def qkv_grad(dqkv=dqkv, x_in=x_in, att=att):
                    dw, db = _wgrad(dqkv, x_in), dqkv.sum(0)
                    for k, lin in enumerate((att.q_linear, att.k_linear, att.v_linear)):
                        oc.accumulate(lin.weight, dw[k * E: (k + 1) * E])
                        oc.accumulate(lin.bias, db[k * E: (k + 1) * E])
                oc.run(qkv_grad, dqkv, x_in)
            d_pre = dx * m_in if m_in is not None else dx
            oc.defer(tokens.view(rows, -1), d_pre, net.features.weight, net.features.bias)
        ctx.saved = None
        return (None, None, None) + (None,) * len(ctx.param_versions)


def forward(net: Any, x: torch.Tensor, oc: Any) -> torch.Tensor:
    """Transformer.forward in training mode inside a Trainer step: features (B, output_size) with autograd history."""
    return _TransformerTrainFn.apply(net, oc, x, *net.parameters())
