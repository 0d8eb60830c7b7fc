"""Accuracy / time of per-layer fp32-promotion periods (env DCSCN_SEG="CNN2=1,CNN5=2,...") on the bench workload:
max |gpu - fp64 oracle| over the first ORACLE_TILES noise tiles and the step time, one fresh engine per setting."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import bench  # noqa: E402
import dcscn_oracle as O  # noqa: E402
from helper import engine as E  # noqa: E402

nt = int(os.environ.get("ORACLE_TILES", "8"))
w = bench.load_weights()
g = torch.Generator().manual_seed(0)
x = (torch.rand(256, 48, 48, 1, generator=g) * 255)
x2 = (torch.rand(256, 96, 96, 1, generator=g) * 255)
y64 = O.Oracle(O.OracleConfig(), w, torch.float64).forward(x[:nt].numpy().astype(np.float64), x2[:nt].numpy().astype(np.float64))
xd, x2d = x.cuda(), x2.cuda()
y = torch.empty_like(x2d)
for setting in sys.argv[1:] or [""]:
    os.environ["DCSCN_SEG"] = setting
    eng = E.Engine(E.make_config())
    eng.set_params(w)
    for _ in range(3):
        eng.forward(xd, x2d, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward(xd, x2d, y)
    e1.record()
    torch.cuda.synchronize()
    err = np.abs(y[:nt].cpu().numpy() - y64)
    per_tile = err.reshape(nt, -1).max(axis=1)
    print("%-60s step %.3f ms  max|gpu-fp64| %.3e  per tile %s" % (setting or "(default)", e0.elapsed_time(e1) / 10, err.max(),
                                                                   " ".join("%.1e" % v for v in per_tile)), flush=True)
    eng.close()
