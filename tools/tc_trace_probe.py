import os, sys, json, torch, numpy as np
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from bcnf_b200 import CondRealNVP_v2
prec = sys.argv[1]
cfg = bench.load_run_config("trajectory_FC_large")
model = bench.build_model(cfg, torch.device("cuda:0"), prec)
flow = model._flow()
with torch.no_grad():
    cond = torch.randn(200, 30, 3).cuda()
    P = flow.project(model.features(cond))
    z = torch.randn(500*200, 19).cuda()
    for _ in range(2): flow.run(True, z, P, inst_period=200)
    torch.cuda.synchronize()
    os.environ["BCNF_TC_TRACE"] = f"gpurun_out/trace_{prec}.txt"
    flow.run(True, z, P, inst_period=200)
    torch.cuda.synchronize()
t = np.loadtxt(f"gpurun_out/trace_{prec}.txt")
t0 = t[0,0]
print("layer | a_ready->wfull | mma issue | acc_full seen(after issue done) | epi own | epi bar | next a_ready wait")
for i in range(8):
    r = t[i]
    nxt = t[i+1,0]
    print(i, int(r[1]-r[0]), int(r[2]-r[1]), int(r[3]-r[2]), int(r[4]-r[3]), int(r[5]-r[4]), int(nxt-r[5]), " | layer period", int(nxt-r[0]))

