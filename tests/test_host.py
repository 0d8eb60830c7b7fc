"""Host-side logic of the drop-in (flags, model-name grammar, utilities, loader) - CPU only."""
import glob
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from helper import args as A
from helper import loader, utilty as util

from conftest import GOLDEN, PKG, ROOT
import dcscn_oracle as O


def fresh_flags(argv):
    f = A._Flags()
    for name, (kind, default, help_text) in A.FLAGS._defs.items():
        f._define(name, default, help_text, kind)
    rest = f.parse(["prog"] + argv)
    return f, rest


def test_flag_defaults_match_reference():
    f, _ = fresh_flags([])
    # helper/args.py:16-98 defaults of the reference
    expect = dict(scale=2, layers=12, filters=196, min_filters=48, filters_decay_gamma=1.5, use_nin=True,
                  nin_filters=64, nin_filters2=32, cnn_size=3, reconstruct_layers=1, reconstruct_filters=32,
                  dropout_rate=0.8, activator="prelu", pixel_shuffler=True, pixel_shuffler_filters=0,
                  self_ensemble=8, batch_norm=False, depthwise_separable=False, clipping_norm=5.0,
                  initializer="he", l2_decay=0.0001, optimizer="adam", beta1=0.9, beta2=0.999, epsilon=1e-8,
                  batch_num=20, batch_image_size=48, training_images=24000, initial_lr=0.002, lr_decay=0.5,
                  lr_decay_epoch=9, end_lr=2e-5, dataset="bsd200", test_dataset="set5", max_value=255.0,
                  channels=1, psnr_calc_border_size=-1, checkpoint_dir="models", output_dir="output",
                  gpu_device_id=0)
    for k, v in expect.items():
        assert getattr(f, k) == v, k


def test_flag_parsing_forms():
    f, rest = fresh_flags(["--scale=4", "--layers", "7", "--nouse_nin", "--batch_norm", "--pixel_shuffler=false",
                           "--filters_decay_gamma=1.2", "positional"])
    assert f.scale == 4 and f.layers == 7 and f.use_nin is False and f.batch_norm is True
    assert f.pixel_shuffler is False and abs(f.filters_decay_gamma - 1.2) < 1e-12
    assert rest == ["prog", "positional"]
    with pytest.raises(SystemExit):
        fresh_flags(["--no_such_flag=1"])


class _NameOnly:
    """get_model_name without constructing an engine."""
    from DCSCN import SuperResolution as _SR
    get_model_name = _SR.get_model_name


def name_for(**kw):
    f, _ = fresh_flags(["--%s=%s" % (k, v) for k, v in kw.items()])
    o = _NameOnly()
    o.layers, o.filters, o.min_filters = f.layers, f.filters, min(f.filters, f.min_filters)
    o.filters_decay_gamma, o.cnn_size, o.scale, o.use_nin = f.filters_decay_gamma, f.cnn_size, f.scale, f.use_nin
    o.nin_filters, o.nin_filters2, o.pixel_shuffler, o.max_value = f.nin_filters, f.nin_filters2, f.pixel_shuffler, f.max_value
    o.activator, o.batch_norm, o.depthwise_separable = f.activator, f.batch_norm, f.depthwise_separable
    o.reconstruct_layers, o.reconstruct_filters = max(f.reconstruct_layers, 1), f.reconstruct_filters
    return o.get_model_name("")


def test_model_name_grammar_matches_shipped_checkpoints():
    # every name below is a file the reference ships under models/
    assert name_for() == "dcscn_L12_F196to48_NIN_A64_PS_R1F32"
    assert name_for(scale=4) == "dcscn_L12_F196to48_Sc4_NIN_A64_PS_R1F32"
    assert name_for(layers=8, filters=96, scale=3) == "dcscn_L8_F96to48_Sc3_NIN_A64_PS_R1F32"
    c = dict(layers=7, filters=32, min_filters=8, filters_decay_gamma=1.2, nin_filters=24, nin_filters2=8,
             reconstruct_layers=0, pixel_shuffler_filters=1)
    assert name_for(**c) == "dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32"
    assert name_for(scale=4, depthwise_separable="true", **c) == "dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32"
    for n in (name_for(), name_for(scale=4), name_for(**c)):
        assert os.path.isfile(os.path.join(GOLDEN, "models", n + ".ckpt.index"))


def test_flip_round_trip_and_shapes():
    g = np.random.RandomState(0)
    img = g.rand(5, 7, 1)
    for t in range(8):
        f = util.flip(img, t)
        assert f.shape == ((7, 5, 1) if t >= 4 else (5, 7, 1))
        np.testing.assert_array_equal(util.flip(f, t, invert=True), img)
        np.testing.assert_array_equal(f, O.flip(img, t))


def test_colour_and_psnr():
    g = np.random.RandomState(1)
    rgb = (g.rand(6, 8, 3) * 255)
    y = util.convert_rgb_to_y(rgb)
    ycc = util.convert_rgb_to_ycbcr(rgb)
    np.testing.assert_allclose(y[:, :, 0], ycc[:, :, 0], atol=1e-9)
    back = util.convert_ycbcr_to_rgb(ycc)
    assert np.abs(back - rgb).max() < 0.5  # the reference's 3-decimal matrices are not exact inverses
    a = g.rand(20, 20, 1) * 255
    b = a + g.randn(20, 20, 1) * 3
    psnr, ssim = util.compute_psnr_and_ssim(a, b, border_size=2)
    assert abs(psnr - O.compute_psnr(a, b, border_size=2)) < 1e-12
    assert 0 < ssim <= 1
    assert util.get_psnr(4.0) == pytest.approx(20 * np.log10(255 / 2.0))


def test_alignment_and_resize():
    img = np.zeros((11, 14, 3), np.uint8)
    assert util.set_image_alignment(img, 4).shape == (8, 12, 3)
    f = np.random.RandomState(0).rand(12, 16, 1) * 255
    up = util.resize_image_by_pil(f, 2)
    assert up.shape == (24, 32, 1) and up.dtype == np.float32
    np.testing.assert_array_equal(up, O.resize_image_by_pil(f, 2))


def test_evaluation_inputs_match_oracle_pipeline():
    f = sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png")))[1]
    true_image = util.set_image_alignment(util.load_image(f, print_console=False), 2)
    lr = loader.build_input_image(true_image, channels=1, scale=2, alignment=2, convert_ycbcr=True)
    o_lr, o_bic, o_true = O.build_inputs_for_evaluate(f, 2)
    np.testing.assert_array_equal(lr, o_lr)
    np.testing.assert_array_equal(util.resize_image_by_pil(lr, 2), o_bic)
    np.testing.assert_array_equal(util.convert_rgb_to_y(true_image), o_true)


def test_split_images():
    img = np.arange(10 * 12, dtype=np.float64).reshape(10, 12, 1)
    w = util.get_split_images(img, 4, stride=2)
    assert w.shape == (4 * 5, 4, 4, 1)
    np.testing.assert_array_equal(w[1, :, :, 0], img[0:4, 2:6, 0])
    assert util.get_split_images(img, 16) is None


def test_sharded_patch_order_is_a_partition():
    """Data-parallel training: every rank draws the same permutation and serves its own residue class, so the ranks'
    patches are disjoint and together cover a pass (helper/loader.py: set_shard)."""
    from helper import loader
    world, count = 4, 37
    served = []
    for rank in range(world):
        ds = loader._ShuffledOrder()
        ds.count = count
        ds.set_shard(rank, world, seed=99)
        ds.init_batch_index()
        got = []
        while ds.index < ds.count:
            got.append(ds.get_next_image_no())
        served.append(got)
    flat = [k for g in served for k in g]
    assert sorted(flat) == list(range(count))
    assert max(len(g) for g in served) - min(len(g) for g in served) <= 1
    # a plain (unsharded) order is unchanged: one pass serves every index once
    ds = loader._ShuffledOrder()
    ds.count = count
    ds.init_batch_index()
    assert sorted(ds.get_next_image_no() for _ in range(count)) == list(range(count))


def test_adam_step_is_recovered_from_beta2_power():
    """beta1_power = 0.9^(t+1) underflows in float32 after ~980 updates; beta2_power = 0.999^(t+1) does not."""
    import DCSCN
    m = object.__new__(DCSCN.SuperResolution)
    m.beta1, m.beta2 = 0.9, 0.999

    class Reader:
        def __init__(self, t):
            self.v = {"beta1_power": np.float32(0.9) ** np.float32(t + 1), "beta2_power": np.float32(0.999 ** (t + 1))}
        def has_tensor(self, k):
            return k in self.v
        def get_tensor(self, k):
            return np.asarray(self.v[k], np.float32)
    for t in (0, 3, 500, 2000, 40000):
        assert abs(m._adam_step_from_powers(Reader(t)) - t) <= max(1, t // 2000), t
    assert float(Reader(2000).v["beta1_power"]) == 0.0           # what the old beta1-only recovery saw


def test_pil_bicubic_restatement_is_bit_exact():
    """helper/pil_resample.py (the tables the device resampler uses) against Pillow itself, up- and down-scaling."""
    from PIL import Image
    from helper import pil_resample as R
    rs = np.random.RandomState(0)
    for (h, w, s) in [(48, 48, 2), (37, 53, 2), (24, 31, 3), (20, 20, 4), (96, 64, 0.5), (99, 63, 1 / 3), (64, 128, 0.25), (1, 7, 2)]:
        a = rs.rand(h, w) * 255
        nw, nh = int(w * s), int(h * s)
        ref = np.asarray(Image.fromarray(a).resize([nw, nh], resample=Image.BICUBIC))
        got = R.resize_float(a, nw, nh)
        assert ref.dtype == np.float32 and np.array_equal(ref, got), (h, w, s)
    # and through the helper the model uses
    a = rs.rand(19, 23, 1) * 255
    assert np.array_equal(util.resize_image_by_pil(a, 2)[:, :, 0], R.resize_float(a[:, :, 0], 46, 38))


def test_package_does_not_import_oracle():
    """The product path must never route through the oracle."""
    for path in glob.glob(os.path.join(PKG, "**", "*.py"), recursive=True) + glob.glob(os.path.join(PKG, "csrc", "*")):
        if path.endswith((".so", ".o")):
            continue
        text = open(path, errors="replace").read()
        assert not re.search(r"^\s*(import|from)\s+dcscn_oracle", text, re.M), path
        assert "oracle/" not in text, path


def test_single_process_paths_never_import_torch(tmp_path):
    """A plain command line (no torchrun) must not pay for `import torch`: rank / world come out as (0, 1) without
    consulting torch.distributed, and train_batch takes the engine's host-buffer step with the reference's step counter
    as the dropout seed (DCSCN.py:415-425)."""
    import subprocess
    import sys
    code = r'''
import os, sys
sys.path.insert(0, %r)
os.environ.pop("WORLD_SIZE", None)
import numpy as np
import DCSCN
assert DCSCN._dist_rank_world() == (0, 1)
m = object.__new__(DCSCN.SuperResolution)
m.batch_input = [np.full((4, 4, 1), i, np.float64) for i in range(3)]
m.batch_input_bicubic = [np.zeros((8, 8, 1)) for _ in range(3)]
m.batch_true = [np.ones((8, 8, 1)) for _ in range(3)]
m.lr, m.step, m.training_step, m.max_value = 0.002, 7, 0, 255.0
m.training_loss_sum = m.training_psnr_sum = 0.0
calls = []
class FakeEngine:
    def train_step_host(self, x, x2, y, lr, seed, apply_update=True):
        calls.append((x.shape, x.dtype, x.flags["C_CONTIGUOUS"], x2.shape, y.shape, lr, seed, apply_update))
        return 4.0, 4.0
    def train_step_data_parallel(self, *a, **k):
        raise AssertionError("data-parallel step in a single process")
m.engine = FakeEngine()
m.train_batch()
assert calls == [((3, 4, 4, 1), np.dtype("float32"), True, (3, 8, 8, 1), (3, 8, 8, 1), 0.002, 7, True)], calls
assert (m.step, m.training_step, m.training_loss_sum) == (8, 1, 4.0)
assert "torch" not in sys.modules, "torch was imported on a single-process path"
print("ok")
''' % PKG
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=str(tmp_path))
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
