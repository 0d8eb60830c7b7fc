"""Timing sweep of engine options on the bench workload (L12 x2, 256 tiles): prints per-layer ms for each setting
and checks the output against the default setting."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
import bench  # noqa: E402
from helper import engine as E  # noqa: E402

batch = int(os.environ.get("BATCH", "256"))
eng = E.Engine(E.make_config())
eng.set_params(bench.load_weights())
g = torch.Generator().manual_seed(0)
x = (torch.rand(batch, 48, 48, 1, generator=g) * 255).cuda()
x2 = (torch.rand(batch, 96, 96, 1, generator=g) * 255).cuda()
y = torch.empty_like(x2)
settings = [dict(s.split("=") for s in a.split(",")) for a in sys.argv[1:]] or [{}]
ref = None
for st in settings:
    for k, v in st.items():
        eng.set_option(k, int(v))
    try:
        for _ in range(2):
            eng.forward(x, x2, y)
        torch.cuda.synchronize()
        eng.set_option("timing", 1)
        acc = {}
        reps = 3
        for _ in range(reps):
            eng.forward(x, x2, y)
            for n, t in eng.timings():
                acc[n] = acc.get(n, 0.0) + t / reps
        eng.set_option("timing", 0)
        out = y.clone()
        if ref is None:
            ref = out
        diff = (out - ref).abs().max().item()
        total = sum(acc.values())
        print("%-40s total %.3f ms  (%.1f Mpix/s)  maxdiff-vs-first %.2e" % (st, total, batch * 96 * 96 / total / 1e3, diff))
        print("    " + " ".join("%s=%.3f" % (n, t) for n, t in acc.items()))
    except Exception as e:  # noqa: BLE001
        print("%-40s FAILED: %s" % (st, e))
        break
