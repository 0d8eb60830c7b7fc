"""The drop-in command lines on a GPU: `evaluate.py` with the reference's own flags against the golden checkpoints
and Set5 must log the PSNR the oracle pins (README: 37.15 dB for c-DCSCN x2)."""
import os
import re
import subprocess
import sys

import pytest

from conftest import GOLDEN, PKG

pytestmark = pytest.mark.gpu


def test_evaluate_cli_c_dcscn_set5(tmp_path):
    cmd = [sys.executable, os.path.join(PKG, "evaluate.py"), "--scale=2", "--layers=7", "--filters=32", "--min_filters=8",
           "--filters_decay_gamma=1.2", "--nin_filters=24", "--nin_filters2=8", "--reconstruct_layers=0", "--self_ensemble=1",
           "--batch_image_size=32", "--pixel_shuffler_filters=1", "--test_dataset=set5", "--save_results=false",
           "--data_dir=" + os.path.join(GOLDEN, "data"), "--checkpoint_dir=" + os.path.join(GOLDEN, "models"),
           "--log_filename=" + str(tmp_path / "log.txt"), "--tf_log_dir=" + str(tmp_path / "tf_log"),
           "--graph_dir=" + str(tmp_path / "graphs"), "--output_dir=" + str(tmp_path / "out")]
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    log = open(tmp_path / "log.txt").read()
    m = re.search(r"Model Average \[set5\] PSNR:([0-9.]+), SSIM:([0-9.nan]+), Time \(s\): ([0-9.]+)", log)
    assert m, log
    assert abs(float(m.group(1)) - 37.148) <= 0.01


def test_sr_cli_writes_outputs(tmp_path):
    img = os.path.join(GOLDEN, "data", "set5", "img_003.png")
    cmd = [sys.executable, os.path.join(PKG, "sr.py"), "--file=" + img, "--scale=2", "--layers=7", "--filters=32",
           "--min_filters=8", "--filters_decay_gamma=1.2", "--nin_filters=24", "--nin_filters2=8", "--reconstruct_layers=0",
           "--self_ensemble=1", "--pixel_shuffler_filters=1", "--checkpoint_dir=" + os.path.join(GOLDEN, "models"),
           "--log_filename=" + str(tmp_path / "log.txt"), "--tf_log_dir=" + str(tmp_path / "tf_log"),
           "--graph_dir=" + str(tmp_path / "graphs"), "--output_dir=" + str(tmp_path / "out")]
    r = subprocess.run(cmd, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = tmp_path / "out" / "dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32"
    names = sorted(os.listdir(out))
    assert names == ["img_003.png", "img_003_bicubic.png", "img_003_bicubic_y.png", "img_003_result.png", "img_003_result_y.png"]
    from PIL import Image
    assert Image.open(out / "img_003_result.png").size == (512, 512)
