"""Where do the planes of a wide-tile run (option wide_tiles=1: two TMEM buffers) differ from the narrow tiling?"""
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

w = bench.load_weights()
g = torch.Generator().manual_seed(0)
n, h, wd = 2, 48, 48
x = (torch.rand(n, h, wd, 1, generator=g) * 255)
x2 = (torch.rand(n, 2 * h, 2 * wd, 1, generator=g) * 255)
y64, inter = O.Oracle(O.OracleConfig(), w, torch.float64).forward(x.numpy().astype(np.float64), x2.numpy().astype(np.float64),
                                                                  return_intermediates=True)
eng = E.Engine(E.make_config())
eng.set_params(w)
xd, x2d = x.cuda(), x2.cuda()
res = {}
for name, opts in (("narrow", {"wide_tiles": 0}), ("wide", {"wide_tiles": 1}), ("wide_seg1", {"wide_tiles": 1, "seg_chunks": 1}),
                   ("narrow_seg1", {"wide_tiles": 0, "seg_chunks": 1})):
    for k, v in opts.items():
        eng.set_option(k, v)
    y = eng.forward(xd, x2d).cpu().numpy()
    a = eng.get_activation("CNN2", (n, h, wd, 166))
    res[name] = (y, a)
    ref = inter["CNN2"]
    e = np.abs(a - ref)
    print("%-12s out err %.3e | CNN2 err vs fp64: max %.3e mean %.3e  signed mean %.3e" % (name, np.abs(y - y64).max(), e.max(), e.mean(),
                                                                                        (a - ref).mean()))
    eng.set_option("seg_chunks", 0)
d = np.abs(res["wide"][1] - res["narrow"][1])
print("CNN2 wide vs narrow: differing elements %.1f %%, max %.3e" % (100.0 * (d > 0).mean(), d.max()))
pc = d.reshape(-1, 166).max(axis=0)
print("per-channel max diff (x1e6):", " ".join("%d" % round(v * 1e6) for v in pc))
px = d.max(axis=3)[0]
print("per-pixel max diff image 0, rows 0..15 cols 0..31 (x1e6):")
for r in range(16):
    print(" ".join("%3d" % round(v * 1e6) for v in px[r, :32]))
refe = np.abs(res["narrow"][1] - inter["CNN2"]).reshape(-1, 166).max(axis=0)
wide = np.abs(res["wide"][1] - inter["CNN2"]).reshape(-1, 166).max(axis=0)
print("per-channel max err vs fp64 narrow (x1e6):", " ".join("%d" % round(v * 1e6) for v in refe))
print("per-channel max err vs fp64 wide   (x1e6):", " ".join("%d" % round(v * 1e6) for v in wide))
