"""Cross-checks the torch-based oracle's reading of the TF op semantics against the plain-C restatement
(oracle/conv_ref.c) and against hand-computed tiny cases.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

import dcscn_oracle as O
from conftest import ROOT


@pytest.fixture(scope="module")
def cref():
    so = os.path.join(ROOT, "oracle", "libconv_ref.so")
    src = os.path.join(ROOT, "oracle", "conv_ref.c")
    if not os.path.isfile(so) or os.path.getmtime(src) > os.path.getmtime(so):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def fptr(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.mark.parametrize("k,cin,cout,h,w", [(3, 5, 7, 6, 9), (1, 11, 4, 5, 5), (3, 1, 3, 4, 4), (5, 2, 2, 7, 6)])
def test_conv_same_matches_c(cref, k, cin, cout, h, w):
    g = np.random.RandomState(k * 100 + cin)
    x = g.randn(2, h, w, cin).astype(np.float32)
    wt = g.randn(k, k, cin, cout).astype(np.float32)
    b = g.randn(cout).astype(np.float32)
    y = np.empty((2, h, w, cout), np.float32)
    cref.conv2d_same_nhwc(fptr(x), fptr(wt), fptr(b), fptr(y), 2, h, w, cin, cout, k)
    xt = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    yo = O.conv2d_same(xt, wt, torch.float64) + torch.from_numpy(b).double().view(1, -1, 1, 1)
    np.testing.assert_allclose(yo.permute(0, 2, 3, 1).numpy(), y, rtol=1e-5, atol=1e-5)


def test_depth_to_space_is_dcr(cref):
    n, h, w, c, r = 1, 2, 3, 4, 2
    x = np.arange(n * h * w * r * r * c, dtype=np.float32).reshape(n, h, w, r * r * c)
    y = np.empty((n, h * r, w * r, c), np.float32)
    cref.depth_to_space_dcr(fptr(x), fptr(y), n, h, w, c, r)
    yo = O.depth_to_space(torch.from_numpy(x).permute(0, 3, 1, 2), r).permute(0, 2, 3, 1).numpy()
    np.testing.assert_array_equal(yo, y)
    # hand check: output (oy*r+i, ox*r+j, ch) <- input channel (i*r+j)*c + ch
    assert y[0, 1, 0, 2] == x[0, 0, 0, (1 * 2 + 0) * 4 + 2]
    assert y[0, 2, 5, 3] == x[0, 1, 2, (0 * 2 + 1) * 4 + 3]


def test_prelu_and_depthwise(cref):
    g = np.random.RandomState(3)
    x = g.randn(1, 5, 6, 4).astype(np.float32)
    a = g.rand(4).astype(np.float32)
    xc = x.copy()
    cref.prelu_nhwc(fptr(xc), fptr(a), ctypes.c_size_t(30), 4)
    xo = O.prelu(torch.from_numpy(x).permute(0, 3, 1, 2), a, torch.float32).permute(0, 2, 3, 1).numpy()
    np.testing.assert_allclose(xo, xc, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(xc, np.where(x > 0, x, a * x), rtol=1e-6, atol=1e-7)
    wd = g.randn(3, 3, 4, 1).astype(np.float32)
    y = np.empty_like(x)
    cref.depthwise_same_nhwc(fptr(x), fptr(wd), fptr(y), 1, 5, 6, 4, 3)
    yo = O.depthwise_same(torch.from_numpy(x).double().permute(0, 3, 1, 2), wd, torch.float64)
    np.testing.assert_allclose(yo.permute(0, 2, 3, 1).numpy(), y, rtol=1e-5, atol=1e-5)


def test_filter_schedule_known_values():
    assert O.feature_filters(O.OracleConfig()) == [196, 166, 148, 133, 120, 108, 97, 86, 76, 66, 57, 48]
    c = O.OracleConfig(layers=7, filters=32, min_filters=8, filters_decay_gamma=1.2)
    assert O.feature_filters(c) == [32, 26, 22, 18, 14, 11, 8]
    # parameter counts of SURVEY.md section 8(a)
    cfg = O.OracleConfig(scale=4)
    n = 0
    for scope, k, cin, cout, bias, prelu in O.layer_table(cfg):
        n += k * k * cin * cout + (cout if bias else 0) + (cout if prelu else 0)
    assert n == 2087102


def test_training_step_closed_form():
    """loss / clip / TF-Adam restatement on a tiny graph: fp64 autograd vs finite differences and closed forms."""
    cfg = O.OracleConfig(scale=2, layers=2, filters=4, min_filters=3, nin_filters=3, nin_filters2=2)
    w = {k: v.astype(np.float64) for k, v in O.he_init_weights(cfg, seed=5).items()}
    orc = O.Oracle(cfg, w, torch.float64)
    g = np.random.RandomState(0)
    x = g.rand(2, 5, 6, 1) * 255
    x2 = g.rand(2, 10, 12, 1) * 255
    y = g.rand(2, 10, 12, 1) * 255
    mse, loss, grads = orc.loss_and_grads(x, x2, y)
    l2 = sum(np.sum(w[n] ** 2) / 2 for n in orc.l2_weight_names())
    assert loss == pytest.approx(mse + cfg.l2_decay * l2, rel=1e-12)
    # finite difference on two parameters
    for name, idx in (("CNN2/conv_W", (1, 1, 2, 0)), ("A1/prelu/A1_prelu", (1,)), ("B2/conv_B", (0,))):
        eps = 1e-5
        w[name][idx] += eps
        lp = orc.loss_and_grads(x, x2, y)[1]
        w[name][idx] -= 2 * eps
        lm = orc.loss_and_grads(x, x2, y)[1]
        w[name][idx] += eps
        assert grads[name][idx] == pytest.approx((lp - lm) / (2 * eps), rel=2e-4, abs=1e-7)
    clipped, norm = orc.clip_by_global_norm(grads)
    assert norm == pytest.approx(np.sqrt(sum(np.sum(v ** 2) for v in grads.values())))
    cn = np.sqrt(sum(np.sum(v ** 2) for v in clipped.values()))
    assert cn == pytest.approx(min(norm, cfg.clipping_norm), rel=1e-9)
    # first Adam step from zero slots: w -= lr * sqrt(1-b2)/(1-b1) * (1-b1) g / (sqrt((1-b2) g^2) + eps)
    m = {n: np.zeros_like(v) for n, v in w.items()}
    v = {n: np.zeros_like(v_) for n, v_ in w.items()}
    before = {n: a.copy() for n, a in w.items()}
    orc.adam_step(clipped, m, v, step=1, lr=0.002)
    name = "CNN1/conv_W"
    gg = clipped[name]
    lr_t = 0.002 * np.sqrt(1 - 0.999) / (1 - 0.9)
    expect = before[name] - lr_t * (0.1 * gg) / (np.sqrt(0.001 * gg * gg) + 1e-8)
    np.testing.assert_allclose(w[name], expect, rtol=1e-10, atol=1e-12)


def test_depthwise_separable_training_step():
    """The oracle side of the depthwise-separable train-step parity (tests/test_gpu_train.py::test_depthwise_separable_*):
    tf.nn.separable_conv2d layers (tf_graph.py:155-177) under the loss of DCSCN.py:340-357.  fp64 autograd against finite
    differences for a depthwise tap, a pointwise weight, a bias and a PReLU slope (with a dropout mask in place); the dead
    `conv_W` of such layers (created by tf_graph.py:183, appended to self.Weights at :212, never read by the forward)
    enters only through the L2 term, so its gradient is exactly l2_decay * conv_W and the forward does not depend on it."""
    cfg = O.OracleConfig(scale=4, layers=3, filters=6, min_filters=3, nin_filters=5, nin_filters2=3, pixel_shuffler_filters=1,
                         depthwise_separable=True)
    w = {k: v.astype(np.float64) for k, v in O.he_init_weights(cfg, seed=7).items()}
    orc = O.Oracle(cfg, w, torch.float64)
    g = np.random.RandomState(1)
    n, h, wd = 2, 5, 4
    x = g.rand(n, h, wd, 1) * 255
    x2 = g.rand(n, 4 * h, 4 * wd, 1) * 255
    y = g.rand(n, 4 * h, 4 * wd, 1) * 255
    masks = {scope: (g.rand(n, cout, h, wd) < 0.8).astype(np.float64)
             for scope, k, cin, cout, bias, prelu in O.layer_table(cfg) if prelu}
    mse, loss, grads = orc.loss_and_grads(x, x2, y, keep_prob=0.8, masks=masks)
    l2 = sum(np.sum(w[nm] ** 2) / 2 for nm in orc.l2_weight_names())
    assert loss == pytest.approx(mse + cfg.l2_decay * l2, rel=1e-12)
    for name, idx in (("CNN2/depthwise_W", (0, 2, 1, 0)), ("A1/pointwise_W", (0, 0, 4, 2)), ("Up-PS/Up-PS_CNN/conv_B", (3,)),
                      ("B1/prelu/B1_prelu", (1,)), ("R-CNN1/depthwise_W", (1, 1, 0, 0)), ("Up-PS2/Up-PS2_CNN/pointwise_W", (0, 0, 2, 1))):
        eps = 1e-5 * max(1.0, abs(w[name][idx]))
        w[name][idx] += eps
        lp = orc.loss_and_grads(x, x2, y, keep_prob=0.8, masks=masks)[1]
        w[name][idx] -= 2 * eps
        lm = orc.loss_and_grads(x, x2, y, keep_prob=0.8, masks=masks)[1]
        w[name][idx] += eps
        assert grads[name][idx] == pytest.approx((lp - lm) / (2 * eps), rel=5e-4, abs=1e-6), name
    for scope, *_ in O.layer_table(cfg):
        np.testing.assert_allclose(grads[scope + "/conv_W"], cfg.l2_decay * w[scope + "/conv_W"], rtol=1e-12, atol=0)
    y_before = orc.forward(x, x2)
    w["CNN1/conv_W"] = w["CNN1/conv_W"] + 1.0
    assert np.array_equal(O.Oracle(cfg, w, torch.float64).forward(x, x2), y_before)
