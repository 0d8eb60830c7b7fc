"""Same-process A/B of two builds of the C-ABI library on the bench workload: the engines of both builds alternate, so
clock / power drift hits both alike.  usage: python scripts/ab_libs.py <libA.so> <libB.so> [rounds]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
import bench  # noqa: E402
from helper import engine as E  # noqa: E402

paths = [os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2])]
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 6
w = bench.load_weights()
engs = []
for p in paths:
    E._lib = None
    E.load_library(p)
    eng = E.Engine(E.make_config())
    eng.set_params(w)
    eng.set_option("timing", 1)
    engs.append(eng)
g = torch.Generator().manual_seed(0)
x = (torch.rand(256, 48, 48, 1, generator=g) * 255).cuda()
x2 = (torch.rand(256, 96, 96, 1, generator=g) * 255).cuda()
y = torch.empty_like(x2)
acc = [{}, {}]
for eng in engs:
    for _ in range(3):
        eng.forward(x, x2, y)
for r in range(rounds):
    for i, eng in enumerate(engs):
        for _ in range(3):
            eng.forward(x, x2, y)
            for n, t in eng.timings():
                acc[i].setdefault(n, []).append(t)
med = [{n: sorted(v)[len(v) // 2] for n, v in a.items()} for a in acc]
print("A = %s\nB = %s" % tuple(paths))
print("%-8s %8s %8s %7s" % ("layer", "A ms", "B ms", "B/A"))
for n in med[0]:
    print("%-8s %8.4f %8.4f %7.3f" % (n, med[0][n], med[1][n], med[1][n] / med[0][n]))
sa, sb = sum(med[0].values()), sum(med[1].values())
print("%-8s %8.4f %8.4f %7.3f" % ("sum", sa, sb, sb / sa))
