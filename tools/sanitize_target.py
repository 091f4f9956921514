"""Small runs of the three kernel families + the training step for compute-sanitizer (memcheck / racecheck).
    compute-sanitizer --tool memcheck python tools/sanitize_target.py [family ...]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bcnf_b200
from bcnf_b200 import CondRealNVP_v2

want = sys.argv[1:] or ["rowthread", "tiled", "tcgen05", "tcgen05_gen1", "train"]
dev = "cuda:0"


def model(nested, precision, blocks=2, n_cond=12):
    torch.manual_seed(0)
    return CondRealNVP_v2(size=19, nested_sizes=nested, n_blocks=blocks, n_conditions=n_cond,
                          feature_networks=[bcnf_b200.ConcatenateCondition(None, n_cond)], dropout=0.1, act_norm=True,
                          precision=precision).to(dev)


g = torch.Generator().manual_seed(1)
y, h = torch.randn(300, 19, generator=g), torch.randn(300, 12, generator=g)
for fam, nested, prec in (("rowthread", [16, 16], "auto"), ("tiled", [40, 40], "fp32"), ("tcgen05", [144, 144], "bf16x3"),
                          ("tcgen05_gen1", [144, 144], "bf16x3")):
    if fam not in want:
        continue
    if fam == "tcgen05_gen1":
        os.environ["BCNF_FLOW_TC"] = "1"
    m = model(nested, prec).eval()
    with torch.no_grad():
        z = m(y, h, log_det_J=True)
        x = m.inverse(z, h)
        s = m.sample(5, h[:40], outer=True)
        flow = m._flow()
        r = torch.zeros(40, 19, dtype=torch.int32, device=dev)
        flow.sample_ranks(200, flow.project(h[:40].to(dev)), y[:40], r, seed=3, inst_period=40)
    torch.cuda.synchronize()
    os.environ.pop("BCNF_FLOW_TC", None)
    print(fam, flow.kernel, int(flow.info.rows_per_cta), float((x.cpu() - y).abs().max()), tuple(s.shape), int(r.sum()), flush=True)
if "train" in want:
    m = model([64, 64], "auto").train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    tr = bcnf_b200.Trainer(m, opt)
    with torch.enable_grad():
        for _ in range(2):
            out = tr.train_batch(y[:64], h[:64])
    torch.cuda.synchronize()
    print("train", out, flush=True)
