"""kind::f16 UMMA throughput in isolation (dcscn_umma_probe): peak at N = 256 and the cost of one K = 16 slice of the
hi/lo scheme as a function of the layer width.  Usage (GPU box): python scripts/umma_probe.py > gpurun_out/umma_probe.txt"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
from helper import engine as E  # noqa: E402


def probe(lib, group, n, mode, iters=4000, sms=148):
    ms, cyc = ctypes.c_float(), ctypes.c_double()
    if lib.dcscn_umma_probe(0, group, n, mode, iters, ctypes.byref(ms), ctypes.byref(cyc)):
        raise RuntimeError(lib.dcscn_last_error().decode())
    slices = iters * 4
    prods = {0: 3, 1: 3, 2: 1}[mode]                     # products of width n per slice
    clusters = sms // group
    macs = clusters * slices * prods * (128 * group) * n * 16
    return dict(group=group, n=n, mode=mode, ms=round(ms.value, 4), cycles_per_slice=round(cyc.value / slices, 1),
                tflops_issued=round(2 * macs / (ms.value * 1e-3) / 1e12, 1),
                mhz=round(cyc.value / (ms.value * 1e-3) / 1e6, 0))


def main():
    lib = E.load_library()
    out = []
    for n in (256, 128):
        out.append(probe(lib, 2, n, 2))
        out.append(probe(lib, 1, n, 2))
    for n in (32, 48, 64, 80, 96, 112, 128, 160, 192, 256):
        out.append(probe(lib, 2, n, 0))
    for n in (32, 48, 64, 80, 96, 112, 128):
        out.append(probe(lib, 1, n, 0))
        out.append(probe(lib, 1, n, 1))
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
