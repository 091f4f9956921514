"""Record state_dict key -> shape of the four BASELINE run configs, from the live reference.

    python tests/golden/make_state_dict_keys.py      (build container only)
"""
import json
import os
import sys

import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_shim import import_reference, reference_config_path  # noqa: E402

NAMES = ["trajectory_FC_small", "trajectory_FC_large", "trajectory_LSTM_large", "trajectory_TRF_large"]


def main():
    ref = import_reference()
    out = {}
    for name in NAMES:
        cfg = yaml.safe_load(open(reference_config_path(f"old/{name}.yaml")))
        torch.manual_seed(0)
        model = ref.CondRealNVP_v2.from_config(cfg)
        out[name] = {"config": cfg, "n_params": int(model.n_params), "n_layers": len(model.layers),
                     "keys": {k: list(v.shape) for k, v in model.state_dict().items()}}
        print(name, model.n_params)
    json.dump(out, open(os.path.join(HERE, "state_dict_keys.json"), "w"))


if __name__ == "__main__":
    main()
