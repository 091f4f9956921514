"""Generate tests/golden/transformer_grads.npz from the LIVE reference (build container only).

    python tests/golden/make_transformer_grad_golden.py

The reference's Transformer encoder (src/bcnf/models/feature_network.py:183-307) in training mode with its dropouts
set to zero, one forward + backward of the scalar sum(h * w) on seeded trajectories (B, 30, 3): state_dict, input, w,
h and the gradient of every parameter.  tests/test_feature_networks.py loads the state_dict into bcnf_b200's module and
compares its autograd gradients (CPU); the GPU tests then hold the Trainer's hand-written backward
(bcnf_b200/trf_train.py) against that module's autograd -- so the encoder's backward is pinned to the reference too.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_shim import import_reference  # noqa: E402

KWARGS = dict(input_size=3, trf_size=32, n_heads=4, ff_size=48, n_blocks=2, output_size=24, dropout=0.0, trf_dropout=0.0)
BATCH = 12


def main():
    import_reference()
    import bcnf.models.feature_network as ref_fn
    torch.manual_seed(17)
    net = ref_fn.Transformer(**KWARGS).train()
    with torch.no_grad():            # LayerNorm gains / shifts away from their (1, 0) initialisation
        for blk in net.layers:
            for ln in (blk.norm1, blk.norm2):
                ln.weight.uniform_(0.5, 1.5)
                ln.bias.uniform_(-0.3, 0.3)
    g = torch.Generator().manual_seed(18)
    x = torch.randn(BATCH, 30, 3, generator=g)
    w = torch.randn(BATCH, KWARGS["output_size"], generator=g)
    out = {f"sd/{k}": v.detach().clone().numpy() for k, v in net.state_dict().items()}
    h = net(x)
    (h * w).sum().backward()
    out["x"], out["w"], out["h"] = x.numpy(), w.numpy(), h.detach().numpy()
    for k, p in net.named_parameters():
        out[f"grad/{k}"] = p.grad.numpy()
    out["meta"] = np.array(json.dumps({"kwargs": KWARGS, "batch": BATCH, "torch": torch.__version__}))
    path = os.path.join(HERE, "transformer_grads.npz")
    np.savez_compressed(path, **out)
    print("->", os.path.getsize(path) // 1024, "KiB,", len([k for k in out if k.startswith("grad/")]), "gradients")


if __name__ == "__main__":
    main()
