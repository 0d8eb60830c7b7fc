"""
bench.py - headline benchmark of the DCSCN hot path (BASELINE.json: "output Mpixels/sec DCSCN L12 x2").

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one forward pass of DCSCN L12 F196->48 x2 over one batch of 256 synthetic 48x48 Y tiles
(BASELINE.json configs[1]) per GPU; `value` is whole-job output Mpixels/s with inputs resident in HBM, `e2e` is
the same metric through the reference-facing host-buffer call (H2D of x and x2 from pinned memory and D2H of y
inside the timed region).  Weights: the reference's own L12 x2 checkpoint (tests/golden/models fixture).

`--impl reference` times the CPU oracle (the reference's graph restated on torch-CPU fp32; TensorFlow is not
installable in this image) on the box's host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))

MODEL = "dcscn_L12_F196to48_NIN_A64_PS_R1F32"
BATCH, TILE, SCALE = 256, 48, 2
# algorithmic work of the L12 x2 forward, no padding / recompute
FLOP_PER_LR_PX_TOTAL = 3508584.0
MAC_PER_LR_PX_TC = 1754292 - 9 * 196 - 4 * 864  # tensor-core layers: all but CNN1 (cin=1) and R-CNN1 (cout=1, 4 HR px per LR px)


def load_weights():
    from helper import tf_bundle
    r = tf_bundle.BundleReader(os.path.join(ROOT, "tests", "golden", "models", MODEL + ".ckpt"))
    return {k: r.get_tensor(k) for k in r.keys()}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1405.3), d.get("bf16_tflops", 1652.1), d.get("hbm_gbs", 6560.6), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


def ncu_traffic():
    """DRAM bytes (read + write) of the same tcgen05 conv launches of one step, from the committed `ncu --set full`
    capture (profiles/r1d_traffic.json, made from profiles/r1d_conv_tc_ncu_summary.csv); None if it is missing."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1d_traffic.json")) as f:
            t = json.load(f)
        return int(t["dram_bytes_read"] + t["dram_bytes_write"]), t["source"]
    except (OSError, KeyError, ValueError):
        return None, None


def cpu_oracle_rate(seconds_target, tiles_per_call=32):
    """Oracle (torch-CPU fp32) throughput in output Mpixels/s on a bounded sample of the same workload."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dcscn_oracle as O
    orc = O.Oracle(O.OracleConfig(), load_weights(), torch.float32)
    g = np.random.RandomState(0)
    x = (g.rand(tiles_per_call, TILE, TILE, 1) * 255).astype(np.float32)
    x2 = (g.rand(tiles_per_call, SCALE * TILE, SCALE * TILE, 1) * 255).astype(np.float32)
    orc.forward(x[:2], x2[:2])  # warm-up (thread pool, oneDNN primitive cache)
    calls, t0 = 0, time.perf_counter()
    while True:
        orc.forward(x, x2)
        calls += 1
        dt = time.perf_counter() - t0
        if dt >= seconds_target or calls >= 64:
            break
    out_px = calls * tiles_per_call * (SCALE * TILE) ** 2
    return out_px / dt / 1e6, dt, calls * tiles_per_call, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dcscn_oracle as O
    orc = O.Oracle(O.OracleConfig(), load_weights(), torch.float32)
    tiles = 16  # bounded sample of the 256-tile batch per step
    g = np.random.RandomState(0)
    x = (g.rand(tiles, TILE, TILE, 1) * 255).astype(np.float32)
    x2 = (g.rand(tiles, SCALE * TILE, SCALE * TILE, 1) * 255).astype(np.float32)
    for _ in range(warm):
        orc.forward(x, x2)
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.forward(x, x2)
    dt = time.perf_counter() - t0
    val = steps * tiles * (SCALE * TILE) ** 2 / dt / 1e6
    sample = "%d of the 256 48x48 tiles per step, torch-CPU fp32 oracle (TensorFlow not installable)" % tiles
    line = {
        "impl": "reference", "metric": "output Mpixels/sec DCSCN L12 x2", "value": val, "unit": "Mpixels/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "DCSCN L12 F196->48 x2 inference, 48x48 Y tiles, CPU sample of %d tiles/step" % tiles},
        "cpu_baseline": {"value": val, "unit": "Mpixels/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": val, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    from helper import engine as E

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    prec = {"f16x3": E.PRECISION_F16X3, "f16x1": E.PRECISION_F16X1}[args.precision]
    eng = E.Engine(E.make_config(device_id=local_rank, precision=prec))
    eng.set_params(load_weights())

    gen = torch.Generator().manual_seed(0 + rank)
    x_host = (torch.rand(BATCH, TILE, TILE, 1, generator=gen) * 255).pin_memory()
    x2_host = (torch.rand(BATCH, SCALE * TILE, SCALE * TILE, 1, generator=gen) * 255).pin_memory()
    y_host = torch.empty(BATCH, SCALE * TILE, SCALE * TILE, 1).pin_memory()
    x, x2 = x_host.cuda(), x2_host.cuda()
    y = torch.empty_like(x2)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        eng.forward(x, x2, y)
    barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        eng.forward(x, x2, y)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out_px_step = BATCH * (SCALE * TILE) ** 2
    value = world * out_px_step * args.steps / (ms / 1e3) / 1e6

    # ---- end to end through the host-buffer API (pinned host memory in, pinned host memory out) ----
    for _ in range(2):
        eng.forward_host(x_host, x2_host, y_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.forward_host(x_host, x2_host, y_host)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = world * out_px_step * args.steps / e2e_s / 1e6

    # ---- per-launch device times (separate pass; CUDA events on the launching stream) ----
    eng.set_option("timing", 1)
    per = {}
    reps = max(2, min(args.steps, 5))
    for _ in range(reps):
        eng.forward(x, x2, y)
        for name, t in eng.timings():
            per[name] = per.get(name, 0.0) + t / reps
    eng.set_option("timing", 0)
    torch.cuda.synchronize()

    if rank == 0:
        sustained, burst, hbm, how = measured_peaks()
        lr_px = BATCH * TILE * TILE
        tc_ms = sum(t for n, t in per.items() if n not in ("CNN1", "R-CNN1"))
        tc_flops = 2.0 * MAC_PER_LR_PX_TC * lr_px
        achieved = tc_flops / (tc_ms / 1e3) / 1e12
        passes = 3 if args.precision == "f16x3" else 1
        roofline = {
            "bound": "tensor", "kernel": "conv_tc_halo1_kernel / conv_tc_pair_kernel (all %d tcgen05 conv launches of one step)" % (len(per) - 2),
            "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (%s)" % how,
            "mma_passes": passes, "frac_of_issued_mma": achieved * passes / sustained,
            "traffic": ncu_traffic()[0], "traffic_unit": "DRAM bytes (read + write) of the same launches of one step",
            "traffic_source": ncu_traffic()[1],
            "hbm_gbs_model": 24700.0 * lr_px / (ms / args.steps / 1e3) / 1e9,
            "launch_ms": {k: round(v, 4) for k, v in per.items()},
        }
        cpu_val, cpu_dt, cpu_tiles, cores = cpu_oracle_rate(args.cpu_seconds)
        line = {
            "metric": "output Mpixels/sec DCSCN L12 x2", "value": value, "unit": "Mpixels/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16x3 (fp16 hi/lo split operands, fp32 accumulate; fp32-equivalent)" if passes == 3 else "f16",
            "data": "synthetic",
            "config": {"workload": "DCSCN L12 F196->48 x2 inference, batch=256 synthetic 48x48 Y-tiles per GPU "
                                   "(BASELINE.json configs[1]), weights = reference L12 x2 checkpoint",
                       "global_batch": BATCH * world, "parallelism": "dp%d (independent tiles, no collective)" % world,
                       "l2": "per-step working set (activation planes) 4.3 GB >> 126 MB L2; no explicit flush"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpixels/s",
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + x2_host.numel() * 4),
                    "d2h_bytes_per_step": int(y_host.numel() * 4)},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": "Mpixels/s", "cores": cores, "kind": "port",
                             "sample": "%d 48x48 tiles of the same workload in %.1f s, torch-CPU fp32 oracle" % (cpu_tiles, cpu_dt)},
            "algorithmic_tflops": FLOP_PER_LR_PX_TOTAL * lr_px * world / (ms / args.steps / 1e3) / 1e12,
        }
        print(json.dumps(line))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def run_secondary(args, rank, world, local_rank):
    """Non-headline workloads of BASELINE.json (`--workload train` = configs[3], `--workload ds` = configs[4]); same JSON
    shape, `config.workload` says which.  One process per GPU; the train step all-reduces one flat gradient buffer."""
    import numpy as np
    import torch
    from helper import engine as E
    from helper import tf_bundle
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def weights(model):
        r = tf_bundle.BundleReader(os.path.join(ROOT, "tests", "golden", "models", model + ".ckpt"))
        return {k: r.get_tensor(k) for k in r.keys()}

    gen = torch.Generator().manual_seed(2 + rank)
    if args.workload == "train":
        per_gpu = 64
        eng = E.Engine(E.make_config(scale=4, device_id=local_rank, dropout_keep=0.8))
        eng.set_params(weights("dcscn_L12_F196to48_Sc4_NIN_A64_PS_R1F32"))
        x = (torch.rand(per_gpu, 48, 48, 1, generator=gen) * 255).cuda()
        x2 = (torch.rand(per_gpu, 192, 192, 1, generator=gen) * 255).cuda()
        y = (torch.rand(per_gpu, 192, 192, 1, generator=gen) * 255).cuda()
        step = lambda i: eng.train_step_data_parallel(x, x2, y, lr=0.002, seed=i * world + rank)
        units, unit, metric = per_gpu * world, "patches/s", "training patches/sec DCSCN L12 x4 (48x48 -> 192x192)"
        name = "DCSCN L12 F196->48 x4 train step, %d 48x48 patches per GPU, dropout keep 0.8, Adam (BASELINE.json configs[3])" % per_gpu
    else:
        eng = E.Engine(E.make_config(scale=4, layers=7, filters=32, min_filters=8, filters_decay_gamma=1.2, nin_filters=24,
                                     nin_filters2=8, reconstruct_layers=0, pixel_shuffler_filters=1, depthwise_separable=True,
                                     device_id=local_rank))
        eng.set_params(weights("dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32"))
        gen = torch.Generator().manual_seed(3 + rank)
        x = (torch.rand(256, 48, 48, 1, generator=gen) * 255).cuda()
        x2 = (torch.rand(256, 192, 192, 1, generator=gen) * 255).cuda()
        yb = torch.empty_like(x2)
        step = lambda i: eng.forward(x, x2, yb)
        units, unit, metric = 256 * 192 * 192 * world / 1e6, "Mpixels/s", "output Mpixels/sec DS c-DCSCN L7 x4"
        name = "depthwise-separable c-DCSCN L7 x4 inference, batch=256 synthetic 48x48 tiles per GPU (BASELINE.json configs[4])"
    for i in range(max(args.warmup, 3)):
        step(i)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.steps):
        step(100 + i)
    ev1.record()
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": metric, "value": units * args.steps / (ms / 1e3), "unit": unit, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16x3 dgrad / fp32 wgrad" if args.workload == "train" else "f32",
            "data": "synthetic", "config": {"workload": name}, "gpu_launches": int(eng.launch_count - l0)}))
    eng.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "f16x1"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0, dest="cpu_seconds")
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "ds"],
                    help="infer = headline (BASELINE configs[1]); train = configs[3]; ds = configs[4]")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # convenience: re-launch ourselves under torchrun, one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if args.workload != "infer":
        run_secondary(args, rank, world, local_rank)
        return
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
