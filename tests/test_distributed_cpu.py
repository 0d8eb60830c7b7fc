"""world_size-2 gloo test (CPU) of the multi-process host logic: the self-ensemble flips sharded over ranks
(DCSCN.SuperResolution.do) must equal the serial ensemble.  The engine call is replaced by the CPU oracle, so
this exercises only the sharding / reduction logic, which is what runs unchanged on N GPUs over NCCL."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _worker(rank, world, port, ensemble, out_dir):
    sys.path.insert(0, PKG)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import DCSCN
    import dcscn_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = O.OracleConfig(scale=2, layers=3, filters=8, min_filters=4, nin_filters=6, nin_filters2=4)
    orc = O.Oracle(cfg, O.he_init_weights(cfg, seed=3), torch.float32)
    m = object.__new__(DCSCN.SuperResolution)
    m.scale, m.self_ensemble, m.max_value, m.resampling_method = 2, ensemble, 255.0, "bicubic"
    m._run = lambda image, bic: orc.forward(
        np.ascontiguousarray(image, np.float32).reshape(1, image.shape[0], image.shape[1], 1),
        np.ascontiguousarray(bic, np.float32).reshape(1, bic.shape[0], bic.shape[1], 1))
    g = np.random.RandomState(0)
    lr = g.rand(9, 13, 1) * 255
    bic = g.rand(18, 26, 1) * 255
    out = m.do(lr, bic)
    np.save(os.path.join(out_dir, "out%d.npy" % rank), out)
    if rank == 0:
        ref = O.do(orc, lr, bic, ensemble)
        np.save(os.path.join(out_dir, "ref.npy"), ref)
    dist.destroy_process_group()


@pytest.mark.parametrize("ensemble", [8, 5])
def test_sharded_ensemble_equals_serial(tmp_path, ensemble):
    port = 29600 + os.getpid() % 300 + ensemble
    mp.spawn(_worker, args=(2, port, ensemble, str(tmp_path)), nprocs=2, join=True)
    ref = np.load(tmp_path / "ref.npy")
    a, b = np.load(tmp_path / "out0.npy"), np.load(tmp_path / "out1.npy")
    np.testing.assert_array_equal(a, b)               # every rank holds the reduced result
    np.testing.assert_allclose(a, ref, rtol=0, atol=1e-9)   # float64 sum in a different order
