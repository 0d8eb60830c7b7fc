"""Same-box A/B of option "graph" (CUDA-graph replay of a forward's launches): batch-256 tiles and batch-1 images.
Usage (GPU box): python scripts/ab_graph.py > gpurun_out/ab_graph.txt"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
from helper import engine as E, tf_bundle  # noqa: E402

MODEL = os.path.join(ROOT, "tests", "golden", "models", "dcscn_L12_F196to48_NIN_A64_PS_R1F32.ckpt")


def timed(eng, x, x2, y, steps):
    for _ in range(3):
        eng.forward(x, x2, y)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.forward(x, x2, y)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    r = tf_bundle.BundleReader(MODEL)
    eng = E.Engine(E.make_config())
    eng.set_params({k: r.get_tensor(k) for k in r.keys()})
    g = torch.Generator().manual_seed(0)
    for n, h, w, steps in ((256, 48, 48, 30), (1, 256, 256, 200), (1, 114, 114, 200), (8, 48, 48, 200)):
        x = (torch.rand(n, h, w, 1, generator=g) * 255).cuda()
        x2 = (torch.rand(n, 2 * h, 2 * w, 1, generator=g) * 255).cuda()
        y = torch.empty_like(x2)
        res = {}
        for rep in range(2):
            for graph in (0, 1):
                eng.set_option("graph", graph)
                res.setdefault(graph, []).append(timed(eng, x, x2, y, steps))
        print(json.dumps(dict(n=n, h=h, w=w, eager_ms=[round(v, 4) for v in res[0]], graph_ms=[round(v, 4) for v in res[1]],
                              replays=eng.graph_replays)))
    eng.close()


if __name__ == "__main__":
    main()
