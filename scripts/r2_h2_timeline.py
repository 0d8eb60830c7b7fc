"""Where the streaming 3x3 kernel's issuing thread and epilogue spend their cycles, per layer of the bench workload.
Needs the diagnostic build:  nvcc ... -DDCSCN_H2_DEBUG -o build_ab/libdcscn_dbg.so engine.cu  and
DCSCN_B200_LIB=build_ab/libdcscn_dbg.so (the product library carries no counters)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
import bench  # noqa: E402
from helper import engine as E  # noqa: E402

eng = E.Engine(E.make_config())
eng.set_params(bench.load_weights())
g = torch.Generator().manual_seed(0)
batch = int(os.environ.get("BATCH", "256"))
x = (torch.rand(batch, 48, 48, 1, generator=g) * 255).cuda()
x2 = (torch.rand(batch, 96, 96, 1, generator=g) * 255).cuda()
y = torch.empty_like(x2)
lib = E.load_library()
lib.dcscn_h2_debug_dump.restype = ctypes.c_int
for _ in range(2):
    eng.forward(x, x2, y)
torch.cuda.synchronize()
lib.dcscn_h2_debug_dump()          # discard the warm-up launches
eng.forward(x, x2, y)
lib.dcscn_h2_debug_dump()
