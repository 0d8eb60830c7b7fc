"""A/B of the 3x3 kernel generations on the bench workload (L12 x2, 256 tiles of 48x48): per-layer device times for each
`halo` setting, agreement between them, and max |gpu - fp64 oracle| on the first tiles of the same noise batch."""
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

batch = int(os.environ.get("BATCH", "256"))
ntile_oracle = int(os.environ.get("ORACLE_TILES", "3"))
w = bench.load_weights()
g = torch.Generator().manual_seed(0)
x = (torch.rand(batch, 48, 48, 1, generator=g) * 255)
x2 = (torch.rand(batch, 96, 96, 1, generator=g) * 255)
y64 = O.Oracle(O.OracleConfig(), w, torch.float64).forward(x[:ntile_oracle].numpy().astype(np.float64),
                                                          x2[:ntile_oracle].numpy().astype(np.float64))
xd, x2d = x.cuda(), x2.cuda()
y = torch.empty_like(x2d)
settings = [dict(s.split("=") for s in a.split(",")) for a in sys.argv[1:]] or [{"halo": "2"}, {"halo": "3"}]
ref = None
eng = E.Engine(E.make_config())
eng.set_params(w)
for st in settings:
    for k, v in st.items():
        eng.set_option(k, int(v))
    for _ in range(3):
        eng.forward(xd, x2d, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng.forward(xd, x2d, y)
    e1.record()
    torch.cuda.synchronize()
    step_ms = e0.elapsed_time(e1) / 10
    eng.set_option("timing", 1)
    acc = {}
    reps = 5
    for _ in range(reps):
        eng.forward(xd, x2d, y)
        for n, t in eng.timings():
            acc.setdefault(n, []).append(t)
    eng.set_option("timing", 0)
    out = y.clone()
    if ref is None:
        ref = out
    diff = (out - ref).abs().max().item()
    err = float(np.abs(out[:ntile_oracle].cpu().numpy() - y64).max())
    med = {n: sorted(v)[len(v) // 2] for n, v in acc.items()}
    print("%-28s step %.3f ms (%.1f Mpix/s)  sum-of-launches %.3f  maxdiff-vs-first %.2e  max|gpu-fp64| (%d tiles) %.3e"
          % (st, step_ms, batch * 96 * 96 / step_ms / 1e3, sum(med.values()), diff, ntile_oracle, err), flush=True)
    print("    " + " ".join("%s=%.3f" % (n, t) for n, t in med.items()), flush=True)
