"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference through oracle/ref_shim.py, builds
CondRealNVP_v2 models on seeded random-init weights (ActNorm perturbed so it is not the
identity), runs forward(log_det_J=True) / inverse / sample on seeded
inputs and stores weights, inputs and outputs.  The fixtures pin oracle/flow_oracle.py
(tests/test_oracle_golden.py) and, on the GPU, the CUDA path (tests/test_gpu_parity.py).
The reference itself cannot travel to the GPU box; these files can.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shim import import_reference, reference_config_path  # noqa: E402

CASES = {
    # the reference's own CPU-runnable BASELINE config, file loaded unchanged
    "fc_small": dict(config_file="old/trajectory_FC_small.yaml", rows=48, cond_shape=(30, 3)),
    # D=21, two-way couplings, every Q identical (random_state set), features passed through
    "d21_two_way": dict(kwargs=dict(size=21, nested_sizes=[32, 32, 32], n_blocks=4, n_conditions=24,
                                    dropout=0.2, act_norm=True, two_way=True, random_state=20240325),
                        rows=33, cond_shape=(24,)),
    # a width that is not a power of two and needs the tiled path
    "h206": dict(kwargs=dict(size=19, nested_sizes=[206, 206, 206], n_blocks=3, n_conditions=40,
                             dropout=0.4, act_norm=True), rows=37, cond_shape=(40,)),
    # the shape of the reference's own unit test (tests/test_cnf.py:18-50): 17x7 inputs,
    # 5 conditions, 5x19 hidden; no ActNorm, dropout 0 (Sequential index stride 2)
    "d7_plain": dict(kwargs=dict(size=7, nested_sizes=[19, 19, 19, 19, 19], n_blocks=3, n_conditions=5,
                                 dropout=0.0, act_norm=False), rows=17, cond_shape=(5,)),
}


def build(ref, case):
    if "config_file" in case:
        cfg = yaml.safe_load(open(reference_config_path(case["config_file"])))
    else:
        c = case["kwargs"]["n_conditions"]
        cfg = {"global": {"parameter_selection": [f"p{i}" for i in range(case["kwargs"]["size"])]},
               "model": {"kwargs": dict(case["kwargs"])},
               "feature_networks": [{"type": "ConcatenateCondition",
                                     "kwargs": {"input_size": None, "output_size": c}}]}
    torch.manual_seed(0)
    model = ref.CondRealNVP_v2.from_config(cfg).eval()
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for layer in model.layers:
            if isinstance(layer, ref.ActNorm):
                layer.scale.copy_(0.5 + torch.rand(layer.scale.shape, generator=g))
                layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g))
    return cfg, model


def main():
    ref = import_reference()
    for name, case in CASES.items():
        cfg, model = build(ref, case)
        d = model.size
        b = case["rows"]
        g = torch.Generator().manual_seed(2)
        y = torch.randn(b, d, generator=g)
        cond = torch.randn(b, *case["cond_shape"], generator=g)
        z_in = torch.randn(b, d, generator=g)
        out = {}
        with torch.no_grad():
            z, h = model(y, cond, log_det_J=True, return_features=True)
            logdet = model.log_det_J.clone()
            x = model.inverse(z_in, cond)
            rt = model.inverse(z, cond)
            # batched sampling, all three _sample modes (cnf.py:564-588), CPU generator seeded
            torch.manual_seed(1234)
            s_outer = model.sample(5, cond[:7], sigma=0.8, outer=True, batch_size=4, sample_batch_size=2)
            torch.manual_seed(1235)
            s_inner = model._sample(7, cond[:7], sigma=1.0, outer=False)
            torch.manual_seed(1236)
            # the all-1-D mode (cnf.py:564-570) needs a condition that is a vector per instance
            s_1d = (model._sample(6, cond[0], sigma=1.0) if cond[0].ndim == 1
                    else torch.zeros(0, d))
            # layer-level calls (the reference's unit test exercises one coupling layer)
            first_c = next(i for i, l in enumerate(model.layers)
                           if isinstance(l, ref.ConditionalAffineCouplingLayer))
            cl = model.layers[first_c]
            cz = cl.forward(y, h, log_det_J=True)
            cld = cl.log_det_J.clone()
            cx = cl.inverse(z_in, h)
            # fp64 truth for error budgeting (log-det accumulated outside the module in fp64)
            import copy
            m64 = copy.deepcopy(model).double()
            y64 = y.double()
            h64 = m64.feature_network_stack(cond.double())
            ld64 = torch.zeros(b, dtype=torch.float64)
            for layer in m64.layers:
                if isinstance(layer, ref.ActNorm):
                    y64 = layer(y64, True)
                else:
                    y64 = layer(y64, h64, True)
                ld64 = ld64 + layer.log_det_J
            # fp64 inverses of the SAME fp32 inputs: the yardstick for ill-conditioned cases
            x64 = m64.inverse(z_in.double(), cond.double())
            rt64 = m64.inverse(z.double(), cond.double())
        for k, v in model.state_dict().items():
            out["sd/" + k] = v.numpy()
        out.update(y=y.numpy(), cond=cond.numpy(), z_in=z_in.numpy(), h=h.numpy(), z=z.numpy(),
                   logdet=logdet.numpy(), x=x.numpy(), roundtrip=rt.numpy(),
                   sample_outer=s_outer.numpy(), sample_inner=s_inner.numpy(), sample_1d=s_1d.numpy(),
                   layer_index=np.int64(first_c), layer_z=cz.numpy(), layer_logdet=cld.numpy(),
                   layer_x=cx.numpy(), z64=y64.numpy(), logdet64=ld64.numpy(),
                   x64=x64.numpy(), roundtrip64=rt64.numpy(), h64=h64.numpy(),
                   )
        # inn_nll_loss from the reference's own utils (utils.py:49-53), via the shim
        import bcnf.utils as ref_utils
        out["nll"] = np.float32(ref_utils.inn_nll_loss(z, logdet).item())
        out["nll_rows"] = ref_utils.inn_nll_loss(z, logdet, reduction="none").numpy()
        out["meta"] = np.array(json.dumps({"config": cfg, "n_params": int(model.n_params),
                                           "n_layers": len(model.layers),
                                           "torch": torch.__version__}))
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "params", model.n_params, "layers", len(model.layers),
              "->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
