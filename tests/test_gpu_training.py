"""Training path (bcnf_b200/train.py + csrc/train_ops.cuh) on the GPU: fused GEMM epilogues vs torch, and
gradient parity of the whole stack against autograd through the CPU oracle (torch back end).

Gradient tolerance: 2e-5 of max|ref grad| per tensor in fp32 (the reference's own fp32 autograd differs
from its fp64 evaluation by ~1e-6 of scale; dropout is tested with the masks the kernels generate,
because no RNG stream can match nn.Dropout's)."""
import numpy as np
import pytest
import torch

import bcnf_b200
from bcnf_b200 import CondRealNVP_v2, _cabi, train
from conftest import rel_err
from oracle import flow_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model(size, nested, n_blocks, n_cond, dropout, two_way=False, seed=0):
    torch.manual_seed(seed)
    model = CondRealNVP_v2(size=size, nested_sizes=nested, n_blocks=n_blocks, n_conditions=n_cond,
                           feature_networks=[bcnf_b200.ConcatenateCondition(None, n_cond)], dropout=dropout,
                           act_norm=True, two_way=two_way)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for layer in model.layers:
            if isinstance(layer, bcnf_b200.ActNorm):
                layer.scale.copy_(0.75 + 0.5 * torch.rand(layer.scale.shape, generator=g))
                layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g))
    return model.to(DEV)


@pytest.mark.parametrize("mode,tol", [(1, 2e-6), (2, 2e-5), (2 | (32 << 4), 2e-5), (2 | (128 << 4), 2e-5)],
                         ids=["fma", "tcgen05", "tcgen05_bn32", "tcgen05_bn128"])
def test_fused_gemm_epilogues_match_torch(mode, tol):
    """fp32 FMA kernel: fp32 rounding.  tcgen05 kernel: 3-pass bf16 split, ~2^-17 per product (stated tolerance 2e-5 of
    max|ref|, fp64 reference)."""
    g = torch.Generator().manual_seed(0)
    B, K, N = 77, 90, 53
    x = torch.randn(B, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / 8).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    close = lambda a, ref: rel_err(a.double().cpu().numpy(), ref.double().cpu().numpy()) < tol
    old = _cabi.lib().bcnf_train_set_gemm_mode(mode)
    try:
        # forward: bias + gelu + dropout, pre-activation saved
        pre, out = torch.empty(B, N, device=DEV), torch.empty(B, N, device=DEV)
        train._gemm(x, (K, 1), w, (1, K), out, B, N, K, epi=_cabi.EPI_BIAS_GELU_DROP, bias=b, save=pre, seed=7, uid=3, p=0.4)
        mask = train.dropout_mask(B, N, 7, 3, 0.4, DEV)
        ref_pre = x.double() @ w.double().t() + b.double()
        assert close(pre, ref_pre)
        assert close(out, torch.nn.functional.gelu(ref_pre) * mask)
        keep = (mask > 0).float().mean().item()
        assert abs(keep - 0.6) < 0.05
        assert all(v == 0.0 or abs(v - 1 / 0.6) < 1e-6 for v in mask.unique().tolist())
        # data gradient: (d W) * gelu'(pre) * mask, with the column sums of the result (bias gradient of the layer below)
        d = torch.randn(B, N, generator=g).to(DEV)
        pre_in = torch.randn(B, K, generator=g).to(DEV)
        din = torch.empty(B, K, device=DEV)
        cs = torch.zeros(K, device=DEV)
        train._gemm(d, (N, 1), w, (K, 1), din, B, K, N, epi=_cabi.EPI_DGELU_DROP, saved=pre_in, seed=7, uid=9, p=0.4, colsum=cs)
        xg = pre_in.double().clone().requires_grad_(True)
        torch.nn.functional.gelu(xg).sum().backward()
        mask_in = train.dropout_mask(B, K, 7, 9, 0.4, DEV)
        ref_din = (d.double() @ w.double()) * xg.grad * mask_in
        assert close(din, ref_din)
        assert close(cs, ref_din.sum(0))
        # weight gradient and bias gradient
        dw = torch.empty(N, K, device=DEV)
        train._gemm(d, (1, N), x, (K, 1), dw, N, K, B)
        assert close(dw, d.double().t() @ x.double())
        db = torch.empty(N, device=DEV)
        train._colsum(d, db)
        assert close(db, d.double().sum(0))
        # beta = 1 accumulates; split-K (tensor-core kernel) reduces through the workspace and leaves it zeroed
        train._gemm(d, (1, N), x, (K, 1), dw, N, K, B, beta=1.0)
        assert close(dw, 2 * (d.double().t() @ x.double()))
        for _ in range(2):
            train._gemm(d, (1, N), x, (K, 1), dw, N, K, B, split_k=0)
            assert close(dw, d.double().t() @ x.double())
        ws, counters = train._workspace(torch.device(DEV))
        assert not ws.any() and not counters.any()
    finally:
        _cabi.lib().bcnf_train_set_gemm_mode(old)


@pytest.mark.parametrize("bn", [32, 64, 128])
@pytest.mark.parametrize("dims", [(77, 53, 90), (256, 526, 526), (300, 130, 1370)], ids=["ragged", "hidden_layer", "projection"])
def test_image_gemm_chain_matches_fp64(dims, bn):
    """TMA-fed tcgen05 GEMM on operand images (bcnf_img_pack -> GEMM with fused epilogue -> c_img -> next GEMM):
    stated tolerance 2e-5 of max|ref| (3-pass bf16 split), fp64 reference."""
    B, N, K = dims
    g = torch.Generator().manual_seed(B + N + K)
    x = torch.randn(B, K, generator=g).to(DEV)
    w1 = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b1 = torch.randn(N, generator=g).to(DEV)
    w2 = (torch.randn(N, N, generator=g) / N ** 0.5).to(DEV)
    dev = torch.device(DEV)
    close = lambda a, ref: rel_err(a.double().cpu().numpy(), ref.double().cpu().numpy()) < 2e-5
    x_img, w1_img, w2t_img = train._Img(dev, B, K), train._Img(dev, N, K), train._Img(dev, N, N)
    h_img = train._Img(dev, B, N)
    # w2t_img: rows = input index of w2, k = output index (the data-gradient orientation)
    train._pack_images([(x, 0, K, 1, B, K, x_img), (w1, 0, K, 1, N, K, w1_img), (w2, 0, 1, N, N, N, w2t_img)], dev)
    old = _cabi.lib().bcnf_train_set_gemm_mode(bn << 4)
    try:
        pre, act = torch.empty(B, N, device=DEV), torch.empty(B, N, device=DEV)
        train._gemm(None, None, None, None, act, B, N, K, epi=_cabi.EPI_BIAS_GELU_DROP, bias=b1, save=pre, seed=11, uid=2, p=0.3,
                    split_k=1, a_img=x_img, b_img=w1_img, c_img=h_img)
        mask = train.dropout_mask(B, N, 11, 2, 0.3, DEV)
        ref_pre = x.double() @ w1.double().t() + b1.double()
        ref_act = torch.nn.functional.gelu(ref_pre) * mask
        assert close(pre, ref_pre) and close(act, ref_act)
        # second GEMM reads the image the first one wrote: d = act . w2 (rows of B = columns of w2), with column sums
        d, cs = torch.empty(B, N, device=DEV), torch.zeros(N, device=DEV)
        train._gemm(None, None, None, None, d, B, N, N, split_k=1, a_img=h_img, b_img=w2t_img, colsum=cs)
        ref_d = ref_act @ w2.double()
        assert close(d, ref_d) and close(cs, ref_d.sum(0))
        # beta = 1
        train._gemm(None, None, None, None, d, B, N, N, split_k=1, a_img=h_img, b_img=w2t_img, beta=1.0)
        assert close(d, 2 * ref_d)
        # weight-gradient mode: dW = act^T x, both images read MN-major (contraction over their rows)
        if bn != 32:
            dw = torch.empty(N, K, device=DEV)
            train._gemm(None, None, None, None, dw, N, K, B, split_k=1, a_img=h_img, b_img=x_img, mn=True)
            ref_dw = ref_act.t() @ x.double()
            assert close(dw, ref_dw)
            train._gemm(None, None, None, None, dw, N, K, B, split_k=1, a_img=h_img, b_img=x_img, mn=True, beta=1.0)
            assert close(dw, 2 * ref_dw)
    finally:
        _cabi.lib().bcnf_train_set_gemm_mode(old)


@pytest.mark.parametrize("shape", [(19, [16] * 3, 4, 24, False, 64), (21, [40, 40], 3, 12, True, 33),
                                   (19, [526] * 5, 2, 1360, False, 256)],
                         ids=["D19_H16", "D21_H40_two_way", "large_conditioner"])
def test_gradients_match_autograd_through_the_oracle(shape):
    size, nested, blocks, n_cond, two_way, rows = shape
    model = _model(size, nested, blocks, n_cond, dropout=0.0, two_way=two_way).train()
    g = torch.Generator().manual_seed(4)
    y = torch.randn(rows, size, generator=g)
    h = torch.randn(rows, n_cond, generator=g)
    y_d, h_d = y.to(DEV).requires_grad_(True), h.to(DEV).requires_grad_(True)
    z = model(y_d, h_d, log_det_J=True)
    loss = bcnf_b200.inn_nll_loss(z, model.log_det_J)
    loss.backward()
    # reference: autograd through the oracle (torch back end) on CPU, same parameters
    sd = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point) for k, v in model.state_dict().items()}
    layers = fo.layers_from_state_dict(sd, convert=lambda v: v)
    y_r, h_r = y.clone().requires_grad_(True), h.clone().requires_grad_(True)
    z_r, ld_r = fo.stack_forward(layers, y_r, h_r)
    loss_r = fo.inn_nll(z_r, ld_r)
    loss_r.backward()
    assert abs(loss.item() - loss_r.item()) < 1e-4 * max(1.0, abs(loss_r.item()))
    assert rel_err(z.detach().cpu().numpy(), z_r.detach().numpy()) < 1e-5
    assert rel_err(y_d.grad.cpu().numpy(), y_r.grad.numpy()) < 2e-5
    assert rel_err(h_d.grad.cpu().numpy(), h_r.grad.numpy()) < 2e-5
    for name, p in model.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        ref = sd[name].grad
        assert ref is not None, name
        assert rel_err(p.grad.cpu().numpy(), ref.numpy()) < 2e-5, name


def test_dropout_gradients_with_the_kernels_own_masks():
    size, nested, n_cond, rows, p, seed = 19, [32, 32], 8, 50, 0.35, 1234
    model = _model(size, nested, 2, n_cond, dropout=p).train()
    g = torch.Generator().manual_seed(5)
    y = torch.randn(rows, size, generator=g).to(DEV).requires_grad_(True)
    h = torch.randn(rows, n_cond, generator=g).to(DEV)
    z, ld = train.stack_forward_train(model, y, h, seed=seed)
    (0.5 * (z ** 2).sum(1) - ld).mean().backward()
    # same computation in plain torch on the GPU with the masks materialised
    y2 = y.detach().clone().requires_grad_(True)
    v, ld2 = y2, torch.zeros(rows, device=DEV)
    params = {n: q.detach().clone().requires_grad_(q.requires_grad) for n, q in model.named_parameters()}
    da = (size + 1) // 2
    for li, layer in enumerate(model.layers):
        pre = f"layers.{li}."
        if isinstance(layer, bcnf_b200.ActNorm):
            v = params[pre + "scale"] * v + params[pre + "bias"]
            ld2 = ld2 + params[pre + "scale"].abs().log().sum()
        elif isinstance(layer, bcnf_b200.OrthonormalTransformation):
            v = v @ params[pre + "orthonormal_matrix"]
        else:
            u = torch.cat([v[:, :da], h], 1)
            idx = [int(k.split(".")[-2]) for k in params if k.startswith(pre + "nn_a.nn.") and k.endswith("weight")]
            for l, j in enumerate(sorted(idx)):
                u = u @ params[f"{pre}nn_a.nn.{j}.weight"].t() + params[f"{pre}nn_a.nn.{j}.bias"]
                if l < len(idx) - 1:
                    u = torch.nn.functional.gelu(u) * train.dropout_mask(rows, u.shape[1], seed, train.layer_uid(li, 0, l), p, DEV)
            t, ls = u[:, : size - da], torch.tanh(u[:, size - da:])
            v = torch.cat([v[:, :da], torch.exp(ls) * v[:, da:] + t], 1)
            ld2 = ld2 + ls.sum(1)
    (0.5 * (v ** 2).sum(1) - ld2).mean().backward()
    assert torch.allclose(z, v, rtol=1e-4, atol=1e-5)
    assert rel_err(y.grad.cpu().numpy(), y2.grad.cpu().numpy()) < 2e-5
    for n, q in model.named_parameters():
        if q.requires_grad:
            assert rel_err(q.grad.cpu().numpy(), params[n].grad.cpu().numpy()) < 2e-5, n
    # a different seed gives a different mask, the same seed the same result
    z3, _ = train.stack_forward_train(model, y.detach(), h, seed=seed)
    z4, _ = train.stack_forward_train(model, y.detach(), h, seed=seed + 1)
    assert torch.equal(z3, z.detach()) and not torch.equal(z4, z.detach())


def test_trainer_step_reduces_the_loss_and_eval_path_sees_the_update():
    import json, os
    from conftest import GOLDEN_DIR
    cfg = json.load(open(os.path.join(GOLDEN_DIR, "state_dict_keys.json")))["trajectory_FC_small"]["config"]
    torch.manual_seed(0)
    model = CondRealNVP_v2.from_config(cfg).to(DEV).train()
    opt = bcnf_b200.OptimizerFactory.get_optimizer("Adam", model.parameters(), {"lr": 2e-4})
    trainer = bcnf_b200.Trainer(model, opt)
    g = torch.Generator().manual_seed(1)
    y = torch.randn(256, 19, generator=g)
    cond = torch.randn(256, 30, 3, generator=g)
    model.eval()
    before = trainer.validate_batch(y, cond)[1]
    model.train()
    losses = [trainer.train_batch(y, cond)[0] for _ in range(30)]
    assert all(np.isfinite(losses))
    model.eval()
    after = trainer.validate_batch(y, cond)[1]
    assert after < before          # 30 Adam steps on one batch lower its NLL; the fused eval kernels saw the new weights


def test_ddp_wraps_the_model_single_rank_nccl():
    import torch.distributed as dist
    import os, socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        model = _model(19, [16] * 2, 3, 8, dropout=0.2).train()
        ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[0])
        opt = torch.optim.Adam(ddp.parameters(), lr=1e-3)
        trainer = bcnf_b200.Trainer(ddp, opt)
        g = torch.Generator().manual_seed(2)
        out = trainer.train_batch(torch.randn(64, 19, generator=g), torch.randn(64, 8, generator=g))
        assert all(np.isfinite(out))
    finally:
        dist.destroy_process_group()


def test_cuda_graph_trainer_matches_eager_statistics():
    # the captured step (zero_grad + forward + backward + Adam) replays with fresh dropout masks and learns
    model = _model(19, [32, 32], 3, 8, dropout=0.3).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, capturable=True)
    trainer = bcnf_b200.Trainer(model, opt, cuda_graph=True)
    g = torch.Generator().manual_seed(9)
    y, c = torch.randn(128, 19, generator=g), torch.randn(128, 8, generator=g)
    losses = [trainer.train_batch(y, c)[0] for _ in range(40)]
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])
    assert len({round(l, 6) for l in losses[:6]}) > 1          # replays are not identical: masks and weights change
    with pytest.raises(ValueError):
        bcnf_b200.Trainer(model, torch.optim.Adam(model.parameters(), lr=1e-3), cuda_graph=True)


def test_two_forwards_before_one_backward_keep_their_own_saved_state():
    """The operand images a forward pass saves for its backward belong to that call (leased pools, bcnf_b200/train.py):
    interleaving two forward passes before the backward gives the same gradients as running them one after the other."""
    model = _model(19, [128, 128, 128], 3, 24, dropout=0.0).train()
    g = torch.Generator().manual_seed(9)
    y1, h1 = torch.randn(200, 19, generator=g).to(DEV), torch.randn(200, 24, generator=g).to(DEV)
    y2, h2 = torch.randn(200, 19, generator=g).to(DEV), torch.randn(200, 24, generator=g).to(DEV)
    nll = lambda z, ld: (0.5 * (z ** 2).sum(1) - ld).mean()
    params = [p for p in model.parameters() if p.requires_grad]
    z1, ld1 = train.stack_forward_train(model, y1, h1, seed=3)
    z2, ld2 = train.stack_forward_train(model, y2, h2, seed=3)          # second forward while the first one's state is alive
    (nll(z1, ld1) + nll(z2, ld2)).backward()
    both = [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    z1, ld1 = train.stack_forward_train(model, y1, h1, seed=3)
    nll(z1, ld1).backward()
    z2, ld2 = train.stack_forward_train(model, y2, h2, seed=3)
    nll(z2, ld2).backward()
    for a, p in zip(both, params):
        assert rel_err(a.cpu().numpy(), p.grad.cpu().numpy()) < 1e-6


def _bf16_round(x32: np.ndarray) -> np.ndarray:
    """fp32 -> nearest bf16 (ties to even), returned as the 16-bit pattern."""
    u = x32.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def _image_reference(x: np.ndarray, rpad: int, chunks: int) -> np.ndarray:
    """The operand image of include/bcnf_b200.h restated in numpy: bf16 hi plane then lo plane, each
    [chunks][rpad rows][128 bytes], 16-byte units of a row XOR-swizzled by (row & 7), zero outside x."""
    rows, k = x.shape
    hi = _bf16_round(x)
    hi_f = (hi.astype(np.uint32) << 16).view(np.float32)
    lo = _bf16_round(x - hi_f)
    planes = np.zeros((2, chunks, rpad, 64), dtype=np.uint16)
    for plane, src in enumerate((hi, lo)):
        for c in range(chunks):
            blk = np.zeros((rows, 64), dtype=np.uint16)
            w = max(0, min(64, k - 64 * c))
            blk[:, :w] = src[:, 64 * c: 64 * c + w]
            units = blk.reshape(rows, 8, 8)                              # 8 units of 8 bf16 (16 bytes)
            r = np.arange(rows)
            for u in range(8):
                planes[plane, c, :rows, :].reshape(rows, 8, 8)[r, u ^ (r & 7)] = units[:, u]
    return planes.reshape(-1).view(np.uint8)


@pytest.mark.parametrize("rows,k", [(5, 3), (77, 90), (256, 526), (300, 1360)])
def test_operand_image_layout_is_bit_exact(rows, k):
    """bcnf_img_pack against the numpy restatement of the format, both source orientations."""
    g = torch.Generator().manual_seed(rows * 1000 + k)
    x = torch.randn(rows, k, generator=g)
    dev = torch.device(DEV)
    im = train._Img(dev, rows, k)
    xd = x.to(DEV)
    train._pack_images([(xd, 0, k, 1, rows, k, im)], dev)
    ref = _image_reference(x.numpy(), im.rpad, im.chunks)
    assert np.array_equal(im.buf.cpu().numpy(), ref)
    # the same matrix read through its transpose (row index contiguous in memory)
    xt = x.t().contiguous().to(DEV)                                      # (k, rows): X(row, col) = xt[col, row]
    im2 = train._Img(dev, rows, k)
    train._pack_images([(xt, 0, 1, rows, rows, k, im2)], dev)
    assert np.array_equal(im2.buf.cpu().numpy(), ref)


def test_gradient_sink_matches_plain_backward_and_graphs_per_shape():
    """Trainer(process_group=...): the stack's gradients are written into views of one flat buffer and all-reduced in
    buckets during the backward (single-rank NCCL group here: the reduction is the identity).  They must equal the
    gradients of the plain autograd path (dropout off), feature-network parameters included; the captured
    step keeps one graph per batch shape and a warm-up that does not train."""
    import torch.distributed as dist
    import os, socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device(DEV))
    try:
        torch.manual_seed(0)
        fc = bcnf_b200.FullyConnectedFeatureNetwork(sizes=[12, 24, 8], dropout=0.0)
        def make():
            torch.manual_seed(3)
            m = CondRealNVP_v2(size=19, nested_sizes=[64, 64], n_blocks=9, n_conditions=8,
                               feature_networks=[bcnf_b200.ConcatenateCondition(None, 12),
                                                 bcnf_b200.FullyConnectedFeatureNetwork(sizes=[12, 24, 8], dropout=0.0)],
                               dropout=0.0, act_norm=True)
            return m.to(DEV).train()
        a, b = make(), make()
        g = torch.Generator().manual_seed(4)
        y, c = torch.randn(96, 19, generator=g), torch.randn(96, 12, generator=g)
        # plain autograd
        ta = bcnf_b200.Trainer(a, torch.optim.SGD(a.parameters(), lr=0.0))
        loss_a, _, _, _ = ta._losses(y, c)
        loss_a.backward()
        # through the sink
        tb = bcnf_b200.Trainer(b, torch.optim.SGD(b.parameters(), lr=0.0), process_group=dist.group.WORLD)
        assert tb._sink is not None and len(tb._other) == 4            # two Linear layers of the feature network
        tb._zero_grad()
        loss_b, _, _, _ = tb._losses(y, c)
        tb._backward(loss_b)
        torch.cuda.synchronize()
        assert torch.equal(loss_a, loss_b)
        for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
            if pa.grad is None:
                assert pb.grad is None or float(pb.grad.abs().max()) == 0.0, n
                continue
            # (bias / ActNorm gradients are accumulated with float atomics: equal up to summation order)
            assert rel_err(pb.grad.cpu().numpy(), pa.grad.cpu().numpy()) < 2e-6, n
        flat = tb._sink.flat
        assert all(p.grad.data_ptr() >= flat.data_ptr() and p.grad.data_ptr() < flat.data_ptr() + 4 * flat.numel()
                   for p in tb._sink.params)                               # .grad really is a view of the flat buffer
        assert len(tb._sink.bucket_bounds(*_units_of(b))) == 4
        # captured steps: one graph per shape, warm-up leaves parameters and Adam state untouched
        m = make()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
        tr = bcnf_b200.Trainer(m, opt, cuda_graph=True, process_group=dist.group.WORLD)
        before = [p.detach().clone() for p in m.parameters()]
        tr.train_batch(y, c)                                               # capture (3 warm-up iterations) + ONE replay
        steps = {int(st["step"].item()) for st in opt.state.values() if "step" in st}
        assert steps == {1}
        tr.train_batch(y[:40], c[:40])                                     # a second shape: its own graph
        tr.train_batch(y, c)
        assert len(tr._graphs) == 2
        assert {int(st["step"].item()) for st in opt.state.values() if "step" in st} == {3}
        assert any(not torch.equal(p, q) for p, q in zip(m.parameters(), before))
        tr.close()
    finally:
        dist.destroy_process_group()


def _units_of(model):
    from bcnf_b200.train import _Spec, _plan, stack_parameters
    kinds = []
    for layer in model.layers:
        kinds.append("actnorm" if isinstance(layer, bcnf_b200.ActNorm) else
                     "ortho" if isinstance(layer, bcnf_b200.OrthonormalTransformation) else "coupling")
    spec = _Spec(kinds, len(model.nested_sizes) + 1, model.two_way, model.size, model.n_conditions, 0.0, 0, None)
    return _plan(spec)[1], stack_parameters(model)


def test_flat_adam_matches_torch_adam():
    """bcnf_adam_flat = torch.optim.Adam's update (trainer.py:271) on the same gradients, incl. weight decay."""
    import copy
    for wd in (0.0, 0.01):
        a = _model(19, [48, 48], 3, 8, dropout=0.0)
        b = copy.deepcopy(a)
        opt_a = bcnf_b200.FlatAdam(a, lr=3e-3, weight_decay=wd)
        opt_b = torch.optim.Adam([p for p in b.parameters() if p.requires_grad], lr=3e-3, weight_decay=wd)
        pa = [p for p in a.parameters() if p.requires_grad]
        pb = [p for p in b.parameters() if p.requires_grad]
        assert all(p.data_ptr() >= opt_a.flat_p.data_ptr() for p in pa)            # parameters are views of the blob
        g = torch.Generator().manual_seed(4)
        for _ in range(6):
            opt_a.zero_grad()
            for x, y in zip(pa, pb):
                gr = torch.randn(x.shape, generator=g).to(DEV) * 0.1
                x.grad.copy_(gr)                                                  # the views stay attached to the blob
                y.grad = gr.clone()
            opt_a.step()
            opt_b.step()
        torch.cuda.synchronize()
        for x, y in zip(pa, pb):
            assert rel_err(x.detach().cpu().numpy(), y.detach().cpu().numpy()) < 2e-6
        assert float(opt_a.step_t) == 6.0
        sd = opt_a.state_dict()
        opt_a.load_state_dict(sd)


def test_fused_nll_matches_inn_nll_loss():
    g = torch.Generator().manual_seed(5)
    for B in (1, 77, 256, 3000):
        z = torch.randn(B, 19, generator=g).to(DEV).requires_grad_()
        ld = torch.randn(B, generator=g).to(DEV).requires_grad_()
        z2, ld2 = z.detach().clone().requires_grad_(), ld.detach().clone().requires_grad_()
        a = bcnf_b200.fused_nll(z, ld)
        b = bcnf_b200.inn_nll_loss(z2, ld2)
        (a * 0.5).backward()
        (b * 0.5).backward()
        assert abs(a.item() - b.item()) <= 2e-6 * max(1.0, abs(b.item()))
        assert rel_err(z.grad.cpu().numpy(), z2.grad.cpu().numpy()) < 1e-6
        assert rel_err(ld.grad.cpu().numpy(), ld2.grad.cpu().numpy()) < 1e-6


def test_trainer_with_flat_adam_follows_torch_adam_and_replays_as_a_graph():
    import copy
    a = _model(19, [48, 48], 3, 8, dropout=0.0).train()
    b = copy.deepcopy(a).train()
    tr_a = bcnf_b200.Trainer(a, bcnf_b200.FlatAdam(a, lr=1e-3))
    tr_b = bcnf_b200.Trainer(b, torch.optim.Adam(b.parameters(), lr=1e-3))
    g = torch.Generator().manual_seed(6)
    y, c = torch.randn(128, 19, generator=g), torch.randn(128, 8, generator=g)
    la = [tr_a.train_batch(y, c)[0] for _ in range(8)]
    lb = [tr_b.train_batch(y, c)[0] for _ in range(8)]
    assert np.allclose(la, lb, rtol=2e-4, atol=2e-4), (la, lb)      # (bias / ActNorm gradients are summed with atomics)
    assert la[-1] < la[0]
    # the eval kernels see the parameters that now live in the optimizer's blob
    a.eval(); b.eval()
    va, vb = tr_a.validate_batch(y, c)[1], tr_b.validate_batch(y, c)[1]
    assert abs(va - vb) < 2e-3 * max(1.0, abs(vb))
    # captured step: learns, and a learning-rate change reaches the replayed kernel
    m = _model(19, [48, 48], 3, 8, dropout=0.2).train()
    opt = bcnf_b200.FlatAdam(m, lr=1e-3)
    tr = bcnf_b200.Trainer(m, opt, cuda_graph=True)
    losses = [tr.train_batch(y, c)[0] for _ in range(30)]
    assert all(np.isfinite(losses)) and np.mean(losses[-5:]) < np.mean(losses[:5])
    assert float(opt.step_t) == 30.0                                   # warm-up steps do not count
    opt.param_groups[0]["lr"] = 0.0
    before = opt.flat_p.clone()
    tr.train_batch(y, c)
    assert torch.equal(before, opt.flat_p)
    tr.close()


def _grad_close(got, ref, tol, name):
    """Relative to max|ref|; a gradient that is identically zero in exact arithmetic (the key bias of an attention layer:
    softmax is invariant to a shift of all scores of a query) is rounding noise on both sides."""
    got, ref = got.detach().cpu().numpy(), ref.detach().cpu().numpy()
    if name.endswith("k_linear.bias"):
        assert np.abs(ref).max() < 1e-5 and np.abs(got).max() < 1e-5, (name, np.abs(ref).max(), np.abs(got).max())
        return
    assert rel_err(got, ref) < tol, (name, rel_err(got, ref))


@pytest.mark.parametrize("kernels", [False, True], ids=["torch_modules", "own_kernels"])
def test_encoder_parameter_gradients_off_the_backward_chain(kernels, monkeypatch):
    """Trainer steps defer the weight / bias / LayerNorm gradients of the Transformer encoder to side streams
    (feature_network.OffChain: the backward chain only carries dx), either around the PyTorch modules
    (BCNF_TRAIN_TRF_KERNELS=0) or around the package's own forward kernels + hand-written backward (bcnf_b200/trf_train.py).
    Same gradients as plain autograd through the modules (dropout off; PyTorch modules: identical arithmetic up to the
    summation order of the bias sums; own kernels: 3-pass bf16 split GEMMs in the forward, stated tolerance 2e-5 of
    max|ref grad|), gradient accumulation over two backward passes, and the captured step trains."""
    from bcnf_b200 import feature_network as fnm, trf_train
    monkeypatch.setattr(trf_train, "ENABLED", kernels)
    tol = 2e-5 if kernels else 2e-6

    def make():
        torch.manual_seed(21)
        m = CondRealNVP_v2(size=19, nested_sizes=[64, 64], n_blocks=3, n_conditions=24,
                           feature_networks=[bcnf_b200.ConcatenateCondition(None, 3),
                                             bcnf_b200.Transformer(input_size=3, trf_size=32, n_heads=4, ff_size=48, n_blocks=2,
                                                                   output_size=24, dropout=0.0, trf_dropout=0.0)],
                           dropout=0.0, act_norm=True)
        with torch.no_grad():
            for blk in m.feature_network_stack.feature_networks[1].layers:
                for ln in (blk.norm1, blk.norm2):
                    ln.weight.uniform_(0.5, 1.5)
                    ln.bias.uniform_(-0.3, 0.3)
        return m.to(DEV).train()
    g = torch.Generator().manual_seed(22)
    y, c = torch.randn(64, 19, generator=g), torch.randn(64, 30, 3, generator=g)
    a, b = make(), make()
    monkeypatch.setattr(train, "_ENC_OFF_CHAIN", False)
    ta = bcnf_b200.Trainer(a, torch.optim.SGD(a.parameters(), lr=0.0))
    loss_a, _, _, _ = ta._losses(y, c)
    ta._backward(loss_a)
    assert getattr(ta, "_oc", None) is None
    monkeypatch.setattr(train, "_ENC_OFF_CHAIN", True)
    tb = bcnf_b200.Trainer(b, torch.optim.SGD(b.parameters(), lr=0.0))
    calls = []
    orig = fnm.OffChain.run
    monkeypatch.setattr(fnm.OffChain, "run", lambda self, fn, *keep: (calls.append(1), orig(self, fn, *keep))[1])
    loss_b, _, _, _ = tb._losses(y, c)
    assert fnm._OFF_CHAIN is None                           # only set while the forward pass runs
    tb._backward(loss_b)
    torch.cuda.synchronize()
    # features, output, and per block: four LayerNorm / Linear groups + (q, k, v as one group | separately) + 2 LayerNorms
    assert len(calls) == (2 + 2 * (4 + 2) if kernels else 2 + 2 * (6 + 2))
    if kernels:
        assert abs(float(loss_a.detach()) - float(loss_b.detach())) < 1e-4 * abs(float(loss_a.detach()))
    else:
        assert torch.equal(loss_a, loss_b)
    for (n, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert (pa.grad is None) == (pb.grad is None), n
        if pa.grad is not None:
            _grad_close(pb.grad, pa.grad, tol, n)
    # a second backward accumulates into the existing .grad tensors
    loss_b2, _, _, _ = tb._losses(y, c)
    tb._backward(loss_b2)
    torch.cuda.synchronize()
    enc_a, enc_b = a.feature_network_stack.feature_networks[1], b.feature_network_stack.feature_networks[1]
    for pa, pb in ((enc_a.output.weight, enc_b.output.weight), (enc_a.layers[0].attention.v_linear.bias, enc_b.layers[0].attention.v_linear.bias),
                   (enc_a.layers[1].norm1.weight, enc_b.layers[1].norm1.weight)):
        assert rel_err(pb.grad.cpu().numpy(), 2.0 * pa.grad.cpu().numpy()) < tol
    # the captured step (forward, backward with the side-stream gradients, Adam) replays and trains
    m = make()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, capturable=True)
    tr = bcnf_b200.Trainer(m, opt, cuda_graph=True)
    first = tr.train_batch(y, c)[0]
    for _ in range(20):
        last = tr.train_batch(y, c)[0]
    assert np.isfinite(last) and last < first
    tr.close()


@pytest.mark.parametrize("pos", [False, True], ids=["plain", "positional"])
def test_transformer_training_kernels_with_dropout_match_autograd(pos, monkeypatch):
    """bcnf_b200/trf_train.py alone, dropout ON: the multiplier tensors it draws are recorded and fed to a functional
    restatement of the module (feature_network.py:183-307) under plain autograd; features and every parameter gradient
    must agree (2e-5 of max|ref|: the forward GEMMs are 3-pass bf16 splits)."""
    import torch.nn.functional as F
    from bcnf_b200 import feature_network as fnm, trf_train
    torch.manual_seed(31)
    net = bcnf_b200.Transformer(input_size=3, trf_size=64, n_heads=4, ff_size=96, n_blocks=3, output_size=40, dropout=0.5,
                                trf_dropout=0.1, add_positional_embeddings=pos).to(DEV).train()
    with torch.no_grad():
        for blk in net.layers:
            for ln in (blk.norm1, blk.norm2):
                ln.weight.uniform_(0.5, 1.5)
                ln.bias.uniform_(-0.3, 0.3)
    B, T, E = 48, 30, 64
    x = torch.randn(B, T, 3, device=DEV)
    w = torch.randn(B, 40, device=DEV)
    masks = []
    orig = trf_train._mask
    monkeypatch.setattr(trf_train, "_mask", lambda shape, p, dev: (masks.append(orig(shape, p, dev)), masks[-1])[1])
    assert trf_train.usable(net, x)
    oc = fnm.OffChain([torch.cuda.Stream(device=DEV), torch.cuda.Stream(device=DEV)])
    h = trf_train.forward(net, x, oc)
    (h * w).sum().backward()
    oc.join(torch.cuda.current_stream(torch.device(DEV)))
    torch.cuda.synchronize()
    got = {n: p.grad.clone() for n, p in net.named_parameters()}
    assert len(masks) == 3 and all(m is not None for m in masks)        # embedding, final state, all blocks at once
    assert abs(float((masks[0] == 0).float().mean()) - 0.5) < 0.02 and abs(float((masks[2] == 0).float().mean()) - 0.1) < 0.02
    masks = masks[:2] + list(masks[2])
    net.zero_grad(set_to_none=True)
    # functional restatement with the same multipliers
    it = iter(masks)
    m_in, m_out = next(it), next(it)
    z = net.features(x) * m_in.view(B, T, E)
    if pos:
        z = z + net._positional(T, z.device)
    for blk in net.layers:
        m1, m2 = next(it), next(it)
        z = blk.norm1(z + blk.attention(z, z, z) * m1.view(B, T, E))
        z = blk.norm2(z + blk.ffn(z) * m2.view(B, T, E))
    h_ref = net.output(z[:, 0, :] * m_out)
    (h_ref * w).sum().backward()
    assert rel_err(h.detach().cpu().numpy(), h_ref.detach().cpu().numpy()) < 2e-5
    for n, p in net.named_parameters():
        _grad_close(got[n], p.grad, 2e-5, n)
