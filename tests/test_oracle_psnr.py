"""Pins the CPU oracle end to end: shipped checkpoints + the restated evaluate pipeline must reproduce the
reference's published PSNR table (README.md:55-65) and the survey's independent per-image values.  CPU only."""
import glob
import json
import os

import numpy as np
import pytest
import torch

import dcscn_oracle as O
from conftest import GOLDEN, MODEL_FLAGS, load_golden_weights

KA = json.load(open(os.path.join(GOLDEN, "psnr_known_answers.json")))


def files(ds):
    return sorted(glob.glob(os.path.join(GOLDEN, "data", ds, "*.png")))


@pytest.mark.parametrize("case", KA["cases"], ids=lambda c: "%s-%s-e%d" % (c["model"][6:24], c["dataset"], c["ensemble"]))
def test_model_psnr(case):
    cfg = O.OracleConfig(**MODEL_FLAGS[case["model"]])
    orc = O.Oracle(cfg, load_golden_weights(case["model"]), torch.float32)
    ps = [O.do_for_evaluate(orc, f, case["ensemble"]) for f in files(case["dataset"])]
    mean = float(np.mean(ps))
    want = case["probe"] if case["probe"] is not None else case["oracle"]   # survey probe, else this oracle's recorded value
    assert abs(mean - want) < 2e-3, (mean, want)
    if case["readme"] is not None:
        # README figures are 2-decimal and their provenance is loose; 0.02 dB covers all rows
        assert abs(mean - case["readme"]) < 0.021, (mean, case["readme"])
    if "per_image" in case:
        np.testing.assert_allclose(ps, case["per_image"], atol=2e-3)


def test_l12_ensemble8_first_image():
    name = "dcscn_L12_F196to48_NIN_A64_PS_R1F32"
    orc = O.Oracle(O.OracleConfig(), load_golden_weights(name), torch.float32)
    p = O.do_for_evaluate(orc, files("set5")[0], 8)
    assert abs(p - KA["l12_x2_set5_ens8_per_image"][0]) < 2e-3, p


@pytest.mark.parametrize("case", KA["bicubic"], ids=lambda c: "%s-x%d" % (c["dataset"], c["scale"]))
def test_bicubic_psnr(case):
    ps = []
    for f in files(case["dataset"]):
        lr, bic, true_y = O.build_inputs_for_evaluate(f, case["scale"])
        ps.append(O.compute_psnr(true_y, bic, border_size=case["scale"]))
    assert abs(np.mean(ps) - case["probe"]) < 2e-3
    assert abs(np.mean(ps) - case["readme"]) < 0.015
