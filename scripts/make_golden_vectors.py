"""
Generates tests/golden/forward_golden.npz: seeded inputs and fp64-oracle outputs of the forward pass for the
shipped checkpoints, so that GPU parity can be checked against committed vectors without running the oracle.
Inputs are 48x48 (and odd-sized) crops of Set5 LR luma with their PIL-bicubic x2 (realistic value ranges).
Run in the build container:  python scripts/make_golden_vectors.py
"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
import dcscn_oracle as O  # noqa: E402
from conftest import MODEL_FLAGS, load_golden_weights  # noqa: E402


def crops(scale, sizes, seed):
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "data", "set5", "*.png")))
    g = np.random.RandomState(seed)
    xs, x2s = [], []
    for i, (h, w) in enumerate(sizes):
        lr, _, _ = O.build_inputs_for_evaluate(files[i % len(files)], scale)
        y0 = g.randint(0, lr.shape[0] - h + 1)
        x0 = g.randint(0, lr.shape[1] - w + 1)
        c = np.ascontiguousarray(lr[y0:y0 + h, x0:x0 + w, :]).astype(np.float32)
        xs.append(c)
        x2s.append(O.resize_image_by_pil(c, scale).astype(np.float32))
    return xs, x2s


def main():
    out = {}
    cases = [
        ("dcscn_L12_F196to48_NIN_A64_PS_R1F32", [(48, 48), (48, 48)], 1),
        ("dcscn_L12_F196to48_NIN_A64_PS_R1F32", [(31, 45)], 2),
        ("dcscn_L12_F196to48_Sc4_NIN_A64_PS_R1F32", [(24, 24)], 3),
        ("dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32", [(48, 48), (48, 48), (48, 48)], 4),
        ("dcscn_L7_F32to8_G1.20_Sc3_NIN_A24_B8_PS_R1F32", [(20, 33)], 5),
        ("dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_R1F32", [(17, 16)], 6),
        ("dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32", [(48, 48), (48, 48)], 7),
        ("dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32", [(13, 70)], 8),
        # round 2: the remaining shipped checkpoints (appended: cases 0..7 stay bit-identical)
        ("dcscn_L12_F196to48_Sc3_NIN_A64_PS_R1F32", [(24, 31)], 9),
        ("dcscn_L8_F96to48_NIN_A64_PS_R1F32", [(48, 48), (48, 48)], 10),
        ("dcscn_L8_F96to48_Sc3_NIN_A64_PS_R1F32", [(21, 40)], 11),
        ("dcscn_L8_F96to48_Sc4_NIN_A64_PS_R1F32", [(19, 23)], 12),
    ]
    for ci, (model, sizes, seed) in enumerate(cases):
        kw = MODEL_FLAGS[model]
        cfg = O.OracleConfig(**kw)
        xs, x2s = crops(cfg.scale, sizes, seed)
        x = np.stack(xs)
        x2 = np.stack(x2s)
        w = load_golden_weights(model)
        y64 = O.Oracle(cfg, w, torch.float64).forward(x.astype(np.float64), x2.astype(np.float64))
        y32 = O.Oracle(cfg, w, torch.float32).forward(x, x2)
        key = "case%d" % ci
        out[key + "_model"] = np.array(model)
        out[key + "_x"] = x
        out[key + "_x2"] = x2
        out[key + "_y64"] = y64                      # float64 reference
        out[key + "_fp32_dev"] = np.array(np.abs(y32 - y64).max())
        print(key, model, x.shape, "fp32-vs-fp64 max %.3e" % np.abs(y32 - y64).max())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "forward_golden.npz"), **out)


if __name__ == "__main__":
    main()
