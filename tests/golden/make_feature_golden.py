"""Generate tests/golden/feature_networks.npz from the LIVE reference (build container only).

    python tests/golden/make_feature_golden.py

The reference's condition encoders (src/bcnf/models/feature_network.py: FullyConnectedFeatureNetwork :114-145,
LSTMFeatureNetwork :148-178, Transformer :263-307) are instantiated small and seeded, run in eval mode on seeded
trajectories of the BASELINE shape (B, 30, 3), and their state_dict, input and output h are stored.  The LSTM is run
with batch == sequence length (the only case the reference's batch-axis pooling accepts, feature_network.py:168-176).
tests/test_feature_networks.py loads the state_dicts into bcnf_b200's own modules and compares h.
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

CASES = {
    "fc": ("FullyConnectedFeatureNetwork", dict(sizes=[90, 48, 48, 40], dropout=0.3), 16),
    "lstm_mean": ("LSTMFeatureNetwork", dict(input_size=3, hidden_size=12, output_size=40, num_layers=2, dropout=0.1,
                                            bidirectional=True, pooling="mean"), 30),
    "lstm_max": ("LSTMFeatureNetwork", dict(input_size=3, hidden_size=10, output_size=24, num_layers=1,
                                           bidirectional=False, pooling="max"), 30),
    "transformer": ("Transformer", dict(input_size=3, trf_size=32, n_heads=4, ff_size=32, n_blocks=2, output_size=40,
                                        dropout=0.3, trf_dropout=0.1), 16),
    "transformer_pos": ("Transformer", dict(input_size=3, trf_size=16, n_heads=2, ff_size=24, n_blocks=1, output_size=24,
                                            add_positional_embeddings=True), 9),
}


def main():
    import_reference()
    import bcnf.models.feature_network as ref_fn
    out = {}
    meta = {}
    for name, (cls, kwargs, batch) in CASES.items():
        torch.manual_seed(7)
        net = getattr(ref_fn, cls)(**kwargs).eval()
        g = torch.Generator().manual_seed(8)
        x = torch.randn(batch, 30, 3, generator=g)
        with torch.no_grad():
            h = net(x)
        for k, v in net.state_dict().items():
            out[f"{name}/sd/{k}"] = v.numpy()
        out[f"{name}/x"] = x.numpy()
        out[f"{name}/h"] = h.numpy()
        meta[name] = {"class": cls, "kwargs": kwargs, "batch": batch, "h_shape": list(h.shape)}
        print(name, cls, "h", tuple(h.shape))
    out["meta"] = np.array(json.dumps({"cases": meta, "torch": torch.__version__}))
    path = os.path.join(HERE, "feature_networks.npz")
    np.savez_compressed(path, **out)
    print("->", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
