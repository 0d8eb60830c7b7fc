"""Multi-GPU paths over NCCL (needs >= 2 GPUs; skipped otherwise): (a) the 8-flip self-ensemble sharded over ranks with one
all-reduce equals the single-GPU result; (b) a data-parallel train step (batch sharded over ranks, one flat gradient
all-reduce, identical clip + Adam everywhere) equals the single-GPU step on the whole batch."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu

KW = dict(scale=2, layers=3, filters=24, min_filters=16, filters_decay_gamma=1.5, nin_filters=16, nin_filters2=16)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, PKG)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import DCSCN
    import dcscn_oracle as O
    from helper import engine as E
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    cfg = O.OracleConfig(**KW)
    wts = O.he_init_weights(cfg, seed=11)
    eng = E.Engine(E.make_config(device_id=rank, dropout_keep=1.0, **KW))
    eng.set_params(wts)
    g = np.random.RandomState(5)
    # (a) sharded ensemble through the drop-in class
    m = object.__new__(DCSCN.SuperResolution)
    m.scale, m.self_ensemble, m.max_value, m.resampling_method, m.engine = 2, 8, 255.0, "bicubic", eng
    lr_img = g.rand(21, 30, 1) * 255
    bic = g.rand(42, 60, 1) * 255
    out = m.do(lr_img, bic)
    np.save(os.path.join(out_dir, "ens%d.npy" % rank), out)
    # (b) data-parallel train step
    x = (g.rand(4, 12, 12, 1) * 255).astype(np.float32)
    x2 = (g.rand(4, 24, 24, 1) * 255).astype(np.float32)
    y = (g.rand(4, 24, 24, 1) * 255).astype(np.float32)
    loss, mse = eng.train_step_data_parallel(np.ascontiguousarray(x[rank::world]), np.ascontiguousarray(x2[rank::world]),
                                             np.ascontiguousarray(y[rank::world]), lr=0.002, seed=7)
    np.save(os.path.join(out_dir, "w%d.npy" % rank), eng.get_param("CNN2/conv_W"))
    np.save(os.path.join(out_dir, "mse%d.npy" % rank), np.array([mse]))
    eng.close()
    dist.destroy_process_group()


def test_two_rank_ensemble_and_data_parallel_step(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import DCSCN
    import dcscn_oracle as O
    from helper import engine as E
    port = 29700 + os.getpid() % 200
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)

    cfg = O.OracleConfig(**KW)
    wts = O.he_init_weights(cfg, seed=11)
    eng = E.Engine(E.make_config(dropout_keep=1.0, **KW))
    eng.set_params(wts)
    g = np.random.RandomState(5)
    m = object.__new__(DCSCN.SuperResolution)
    m.scale, m.self_ensemble, m.max_value, m.resampling_method, m.engine = 2, 8, 255.0, "bicubic", eng
    lr_img = g.rand(21, 30, 1) * 255
    bic = g.rand(42, 60, 1) * 255
    single = m.do(lr_img, bic)
    e0, e1 = np.load(tmp_path / "ens0.npy"), np.load(tmp_path / "ens1.npy")
    np.testing.assert_array_equal(e0, e1)
    np.testing.assert_allclose(e0, single, rtol=0, atol=1e-9)      # same 8 forward results, float64 sum order differs

    x = (g.rand(4, 12, 12, 1) * 255).astype(np.float32)
    x2 = (g.rand(4, 24, 24, 1) * 255).astype(np.float32)
    y = (g.rand(4, 24, 24, 1) * 255).astype(np.float32)
    loss, mse = eng.train_step_host(x, x2, y, lr=0.002, seed=7)
    w_single = eng.get_param("CNN2/conv_W")
    w0, w1 = np.load(tmp_path / "w0.npy"), np.load(tmp_path / "w1.npy")
    np.testing.assert_array_equal(w0, w1)                          # replicas stay identical
    assert np.abs(w0 - w_single).max() <= 0.05 * 0.002             # Adam step ~lr*sign(g), see test_gpu_train.py
    assert float(np.load(tmp_path / "mse0.npy")[0]) == pytest.approx(mse, rel=1e-4)
    eng.close()
