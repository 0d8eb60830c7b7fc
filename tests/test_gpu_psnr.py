"""PSNR / pixel parity on Set5 through the drop-in class (DCSCN.SuperResolution + evaluate pipeline), GPU."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import dcscn_oracle as O
from conftest import GOLDEN, MODEL_FLAGS, load_golden_weights

pytestmark = pytest.mark.gpu

KA = json.load(open(os.path.join(GOLDEN, "psnr_known_answers.json")))


def build_model(tmp_path, flag_args, ensemble):
    from helper import args as A
    import DCSCN
    f = A._Flags()
    for name, (kind, default, help_text) in A.FLAGS._defs.items():
        f._define(name, default, help_text, kind)
    f.parse(["prog", "--checkpoint_dir=" + os.path.join(GOLDEN, "models"), "--self_ensemble=%d" % ensemble,
             "--log_filename=" + str(tmp_path / "log.txt"), "--tf_log_dir=" + str(tmp_path / "tf_log"),
             "--graph_dir=" + str(tmp_path / "graphs"), "--output_dir=" + str(tmp_path / "out")] + flag_args)
    m = DCSCN.SuperResolution(f, model_name=f.model_name)
    m.build_graph()
    m.build_summary_saver()
    m.init_all_variables()
    m.load_model(f.load_model_name)
    return m


CD = ["--scale=2", "--layers=7", "--filters=32", "--min_filters=8", "--filters_decay_gamma=1.2", "--nin_filters=24",
      "--nin_filters2=8", "--reconstruct_layers=0", "--pixel_shuffler_filters=1"]


@pytest.mark.parametrize("flag_args,model,ens", [
    (CD, "dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32", 1),
    ([], "dcscn_L12_F196to48_NIN_A64_PS_R1F32", 1),
    (["--scale=4", "--depthwise_separable=true"] + CD[1:], "dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32", 1),
], ids=["c-DCSCN", "L12", "DS-x4"])
def test_set5_psnr_and_pixels(tmp_path, flag_args, model, ens):
    m = build_model(tmp_path, flag_args, ens)
    assert m.name == model
    case = [c for c in KA["cases"] if c["model"] == model and c["dataset"] == "set5" and c["ensemble"] == ens][0]
    orc = O.Oracle(O.OracleConfig(**MODEL_FLAGS[model]), load_golden_weights(model), torch.float32)
    orc64 = O.Oracle(O.OracleConfig(**MODEL_FLAGS[model]), load_golden_weights(model), torch.float64)
    ps = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png"))):
        psnr, ssim = m.do_for_evaluate(f)
        p_orc = O.do_for_evaluate(orc, f, ens)
        assert abs(psnr - p_orc) <= 0.01, (f, psnr, p_orc)       # north_star: PSNR within 0.01 dB
        ps.append(psnr)
        lr, bic, _ = O.build_inputs_for_evaluate(f, m.scale)
        out = m.do(lr, bic)
        ref32 = O.do(orc, lr, bic, ens)
        ref64 = O.do(orc64, lr.astype(np.float64), bic.astype(np.float64), ens)
        # north_star: 1e-3 absolute (fp32).  The fp32 CPU forward itself sits up to ~9e-4 from the exact (fp64) result on
        # these images, so the bar is applied where it is meaningful: distance to the exact result <= 1e-3, and distance
        # to the fp32 CPU forward bounded by 1e-3 plus that forward's own rounding error.
        assert np.abs(out - ref64).max() <= 1e-3, f
        assert np.abs(out - ref32).max() <= 1e-3 + np.abs(ref32 - ref64).max(), f
    assert abs(np.mean(ps) - case["probe"]) <= 0.01
    if case["readme"] is not None:
        assert abs(np.mean(ps) - case["readme"]) <= 0.021


@pytest.mark.parametrize("flag_args,model", [
    (CD, "dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32"),
    ([], "dcscn_L12_F196to48_NIN_A64_PS_R1F32"),
], ids=["c-DCSCN", "L12"])
def test_set14_psnr(tmp_path, flag_args, model):
    """north_star's PSNR gate names Set5 AND Set14.  Set14 holds the grayscale img_003 (mode L: the uint8-through-PIL
    branch of DCSCN.py:688-696) and non-square images up to 360x250 LR.  Every image's PSNR through the drop-in class is
    held to the CPU oracle's (c-DCSCN) or to the survey's per-image known answers (L12, 3 decimals) within 0.01 dB."""
    m = build_model(tmp_path, flag_args, 1)
    case = [c for c in KA["cases"] if c["model"] == model and c["dataset"] == "set14" and c["ensemble"] == 1][0]
    files = sorted(glob.glob(os.path.join(GOLDEN, "data", "set14", "*.png")))
    assert len(files) == 14
    orc = O.Oracle(O.OracleConfig(**MODEL_FLAGS[model]), load_golden_weights(model), torch.float32)
    ps = []
    for i, f in enumerate(files):
        psnr, _ = m.do_for_evaluate(f)
        ps.append(psnr)
        if "per_image" in case:
            assert abs(psnr - case["per_image"][i]) <= 0.01 + 5e-4, (f, psnr, case["per_image"][i])   # known answers carry 3 decimals
        else:
            assert abs(psnr - O.do_for_evaluate(orc, f, 1)) <= 0.01, f
    gray = files[2]
    lr, bic, _ = O.build_inputs_for_evaluate(gray, 2)
    assert lr.shape[2] == 1 and float(np.abs(lr - np.rint(lr)).max()) == 0.0      # the monochrome branch really is integer-valued
    ref = O.do(orc, lr, bic, 1)
    assert np.abs(m.do(lr, bic) - ref).max() <= 1.5e-3
    assert abs(np.mean(ps) - case["probe"]) <= 0.01
    assert abs(np.mean(ps) - case["readme"]) <= 0.021


@pytest.mark.parametrize("flag_args,model", [
    (["--scale=3"], "dcscn_L12_F196to48_Sc3_NIN_A64_PS_R1F32"),
    (["--layers=8", "--filters=96"], "dcscn_L8_F96to48_NIN_A64_PS_R1F32"),
    (["--layers=8", "--filters=96", "--scale=3"], "dcscn_L8_F96to48_Sc3_NIN_A64_PS_R1F32"),
    (["--layers=8", "--filters=96", "--scale=4"], "dcscn_L8_F96to48_Sc4_NIN_A64_PS_R1F32"),
], ids=["L12-x3", "L8-x2", "L8-x3", "L8-x4"])
def test_remaining_shipped_checkpoints_set5(tmp_path, flag_args, model):
    """the other checkpoints the reference ships (README.md:80,100,132 use --layers=8 --filters=96;
    x3 is a single 3x pixel shuffle with 9 * 96 = 864 Up-PS channels, DCSCN.py:305-308).  Per image: PSNR within 0.01 dB
    of the CPU oracle's recorded value, pixels within 1e-3 of the fp64 oracle on one image."""
    m = build_model(tmp_path, flag_args, 1)
    assert m.name == model
    case = [c for c in KA["cases"] if c["model"] == model and c["ensemble"] == 1][0]
    files = sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png")))
    ps = [m.do_for_evaluate(f)[0] for f in files]
    for p, want in zip(ps, case["per_image"]):
        assert abs(p - want) <= 0.01 + 5e-4, (model, ps, case["per_image"])
    assert abs(np.mean(ps) - case["oracle"]) <= 0.01
    lr, bic, _ = O.build_inputs_for_evaluate(files[3], m.scale)
    orc64 = O.Oracle(O.OracleConfig(**MODEL_FLAGS[model]), load_golden_weights(model), torch.float64)
    ref = O.do(orc64, lr.astype(np.float64), bic.astype(np.float64), 1)
    assert np.abs(m.do(lr, bic) - ref).max() <= 1e-3


def test_l12_x3_set5_ensemble8_matches_the_readme(tmp_path):
    model = "dcscn_L12_F196to48_Sc3_NIN_A64_PS_R1F32"
    m = build_model(tmp_path, ["--scale=3"], 8)
    case = [c for c in KA["cases"] if c["model"] == model and c["ensemble"] == 8][0]
    ps = [m.do_for_evaluate(f)[0] for f in sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png")))]
    assert abs(np.mean(ps) - case["oracle"]) <= 0.01
    assert abs(np.mean(ps) - case["readme"]) <= 0.021          # README.md:58: 34.06


def test_l12_x4_set5_ensemble8_psnr(tmp_path):
    """The x4 flagship (two pixel-shuffler stages) with the default self_ensemble = 8: Set5 average against the survey's
    known answer 31.703 dB (README: 31.72) and, on one image, pixels against the fp64 oracle's ensemble."""
    model = "dcscn_L12_F196to48_Sc4_NIN_A64_PS_R1F32"
    m = build_model(tmp_path, ["--scale=4"], 8)
    assert m.name == model
    case = [c for c in KA["cases"] if c["model"] == model and c["ensemble"] == 8][0]
    files = sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png")))
    ps = [m.do_for_evaluate(f)[0] for f in files]
    assert abs(np.mean(ps) - case["probe"]) <= 0.01
    assert abs(np.mean(ps) - case["readme"]) <= 0.021
    lr, bic, _ = O.build_inputs_for_evaluate(files[4], 4)          # 86x57 LR
    orc64 = O.Oracle(O.OracleConfig(scale=4), load_golden_weights(model), torch.float64)
    ref = O.do(orc64, lr.astype(np.float64), bic.astype(np.float64), 8)
    assert np.abs(m.do(lr, bic) - ref).max() <= 1e-3


def test_self_ensemble_8_matches_oracle(tmp_path):
    m = build_model(tmp_path, [], 8)
    model = "dcscn_L12_F196to48_NIN_A64_PS_R1F32"
    orc = O.Oracle(O.OracleConfig(), load_golden_weights(model), torch.float32)
    f = sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png")))[4]  # 172x114 LR: both orientations
    lr, bic, true_y = O.build_inputs_for_evaluate(f, 2)
    out = m.do(lr, bic)
    ref = O.do(orc, lr, bic, 8)
    assert np.abs(out - ref).max() <= 1.5e-3   # fp32 CPU forward: its own rounding error is ~5e-4 here (see above)
    assert abs(O.compute_psnr(true_y, out, 2) - KA["l12_x2_set5_ens8_per_image"][4]) <= 0.01


@pytest.mark.parametrize("flips", [2, 5, 8])
def test_device_ensemble_equals_the_serial_flip_loop(tmp_path, flips):
    """`dcscn_forward_ensemble` (flips, two batched forwards and the float64 mean on the GPU) against the reference's
    serial loop `output += flip(run(flip(x, i)), i, invert=True)` (DCSCN.py:560-575) over the same engine.  The batched
    forward may cut K into different promotion segments than the n = 1 forward (the segment rule looks at the tile
    count), so the two agree to fp32 rounding of a 0..255 pixel, not bit for bit: 2e-4."""
    from helper import utilty as util
    m = build_model(tmp_path, CD, flips)
    f = sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png")))[4]      # non-square: both orientations
    lr, bic, _ = O.build_inputs_for_evaluate(f, 2)
    out = m.do(lr, bic)
    ref = np.zeros_like(out)
    for i in range(flips):
        c = lambda a: np.ascontiguousarray(a[None], dtype=np.float32)
        y = m.engine.forward_host(c(util.flip(lr, i)), c(util.flip(bic, i)))
        ref += util.flip(y[0], i, invert=True)
    ref /= flips
    assert out.dtype == np.float64 and out.shape == ref.shape
    assert np.abs(out - ref).max() <= 2e-4, float(np.abs(out - ref).max())


def test_save_and_reload_checkpoint(tmp_path):
    """save_model writes a TF V2 bundle that load_model (and the reference's Saver) can read back."""
    import shutil
    m = build_model(tmp_path, CD, 1)
    f = sorted(glob.glob(os.path.join(GOLDEN, "data", "set5", "*.png")))[0]
    p0, _ = m.do_for_evaluate(f)
    m.checkpoint_dir = str(tmp_path / "ckpt")
    m.save_model()
    m.init_all_variables()                       # scramble
    p_rand, _ = m.do_for_evaluate(f)
    assert abs(p_rand - p0) > 1.0
    m.load_model()
    p1, _ = m.do_for_evaluate(f)
    assert p1 == p0


def test_checkpoint_carries_optimizer_state(tmp_path):
    """save_model writes what the reference's tf.train.Saver() writes (trainables, `<var>/Adam`, `<var>/Adam_1`,
    beta1_power, beta2_power); load_model(restore_optimizer=True) brings the Adam state back, so a resumed run takes
    exactly the step the uninterrupted run would have taken."""
    from helper import tf_bundle
    m = build_model(tmp_path, CD, 1)
    g = np.random.RandomState(0)
    x = (g.rand(4, 16, 16, 1) * 255).astype(np.float32)
    x2 = (g.rand(4, 32, 32, 1) * 255).astype(np.float32)
    y = (g.rand(4, 32, 32, 1) * 255).astype(np.float32)
    for i in range(3):
        m.engine.train_step_host(x, x2, y, lr=1e-3, seed=i)
    m.checkpoint_dir = str(tmp_path / "ckpt")
    m.save_model()
    r = tf_bundle.BundleReader(os.path.join(m.checkpoint_dir, m.name + ".ckpt"))
    names = set(r.keys())
    shapes = m.engine.param_shapes()
    assert names == set(shapes) | {v + s for v in shapes for s in ("/Adam", "/Adam_1")} | {"beta1_power", "beta2_power"}
    assert float(r.get_tensor("beta1_power")) == pytest.approx(m.beta1 ** 4, rel=1e-6)
    m.engine.train_step_host(x, x2, y, lr=1e-3, seed=7)          # the uninterrupted run's 4th step
    want = {v: m.engine.get_param(v) for v in shapes}

    m2 = build_model(tmp_path, CD, 1)
    m2.checkpoint_dir = m.checkpoint_dir
    m2.load_model(restore_optimizer=True)
    assert m2.engine.adam_step == 3
    m2.engine.train_step_host(x, x2, y, lr=1e-3, seed=7)
    for v in shapes:
        assert np.abs(m2.engine.get_param(v) - want[v]).max() <= 1e-5, v   # one Adam step moves a weight by <= lr = 1e-3; fp32 atomics reorder a few small sums
