"""
CPU ORACLE for the DCSCN forward / backward hot path.  TEST INFRASTRUCTURE ONLY.

This is a CPU restatement of what the reference's TensorFlow graph computes
(`/root/reference/DCSCN.py:222-332` build_graph, `:334-413` build_optimizer and
`helper/tf_graph.py:77-249`), written with torch-CPU convolutions (fp32, the
reference dtype, or fp64).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import it; the product
path (`dcscn-super-resolution_b200/`) never does.

Parity pinning: the reference has no tests and TensorFlow is not installable
in this image, so the reference itself cannot run here.  The oracle is pinned
instead by (a) the crc32c / value known-answers of the shipped checkpoints,
(b) the README PSNR table (README.md:55-65) reproduced to 2 decimals on
Set5 / Set14 with the shipped `models/*.ckpt` through the restated
`do_for_evaluate` host pipeline (DCSCN.py:672-703) - see
tests/test_oracle_psnr.py and tests/golden/psnr_known_answers.json - and
(c) an independent plain-C direct-convolution restatement (oracle/conv_ref.c)
that cross-checks the torch convolution semantics (SAME padding,
cross-correlation, HWIO filters, DCR depth_to_space).
At the TF-kernel boundary itself parity is "unpinned" (no reference golden
tensors exist); this is stated in DESIGN.md.

TF semantics encoded here:
  * tf.nn.conv2d SAME, stride 1, NHWC activations, HWIO filters, cross-correlation
    (tf_graph.py:105);  bias add (tf_graph.py:109)
  * PReLU  relu(x) + alpha * (x - |x|) * 0.5          (tf_graph.py:89-94)
  * dropout after the activation, identity at keep=1   (tf_graph.py:129-130)
  * H_concat = concat(CNN1..CNNL, axis=3)              (DCSCN.py:258-259)
  * Concat2  = concat([B2, A1], axis=3)                (DCSCN.py:281)
  * Up-PS: conv(+bias, no activation) then tf.depth_to_space (DCR order)
    (tf_graph.py:238-249), x4 = two x2 stages          (DCSCN.py:298-304)
  * R-CNN1: conv without bias/activation, then + x2    (DCSCN.py:318-325)
  * depthwise-separable variant: depthwise kxk (multiplier 1) -> pointwise 1x1
    -> bias -> PReLU                                   (tf_graph.py:155-216)
"""

import math

import numpy as np
import torch
import torch.nn.functional as F


class OracleConfig:
    """The subset of helper/args.py flags that shape the graph."""

    def __init__(self, scale=2, layers=12, filters=196, min_filters=48, filters_decay_gamma=1.5,
                 use_nin=True, nin_filters=64, nin_filters2=32, cnn_size=3, reconstruct_layers=1,
                 reconstruct_filters=32, pixel_shuffler_filters=0, depthwise_separable=False,
                 channels=1, l2_decay=0.0001, clipping_norm=5.0, beta1=0.9, beta2=0.999,
                 epsilon=1e-8):
        self.scale = scale
        self.layers = layers
        self.filters = filters
        self.min_filters = min(filters, min_filters)        # DCSCN.py:37
        self.filters_decay_gamma = filters_decay_gamma
        self.use_nin = use_nin
        self.nin_filters = nin_filters
        self.nin_filters2 = nin_filters2
        self.cnn_size = cnn_size
        self.reconstruct_layers = max(reconstruct_layers, 1)  # DCSCN.py:42
        self.reconstruct_filters = reconstruct_filters
        self.pixel_shuffler_filters = pixel_shuffler_filters
        self.depthwise_separable = depthwise_separable
        self.channels = channels
        self.l2_decay = l2_decay
        self.clipping_norm = clipping_norm
        self.beta1, self.beta2, self.epsilon = beta1, beta2, epsilon


def feature_filters(cfg):
    """Filter-count schedule of the feature-extraction stack (DCSCN.py:240-244)."""
    out = []
    n = cfg.filters
    for i in range(cfg.layers):
        if cfg.min_filters != 0 and i > 0:
            x1 = i / float(cfg.layers - 1)
            y1 = pow(x1, 1.0 / cfg.filters_decay_gamma)
            n = int((cfg.filters - cfg.min_filters) * (1 - y1) + cfg.min_filters)
        out.append(n)
    return out


def layer_table(cfg):
    """
    [(scope, k, cin, cout, use_bias, use_prelu)] in graph-construction order,
    i.e. the order of self.Weights (DCSCN.py:240-323).
    """
    assert cfg.use_nin, "oracle covers the NIN graph (all shipped checkpoints)"
    layers = []
    cin = cfg.channels
    total = 0
    for i, cout in enumerate(feature_filters(cfg)):
        layers.append(("CNN%d" % (i + 1), cfg.cnn_size, cin, cout, True, True))
        cin = cout
        total += cout
    layers.append(("A1", 1, total, cfg.nin_filters, True, True))
    layers.append(("B1", 1, total, cfg.nin_filters2, True, True))
    layers.append(("B2", 3, cfg.nin_filters2, cfg.nin_filters2, True, True))
    cin = cfg.nin_filters + cfg.nin_filters2
    ps_out = cfg.pixel_shuffler_filters if cfg.pixel_shuffler_filters != 0 else cin
    if cfg.scale == 4:
        layers.append(("Up-PS/Up-PS_CNN", cfg.cnn_size, cin, 4 * cin, True, False))
        layers.append(("Up-PS2/Up-PS2_CNN", cfg.cnn_size, cin, 4 * ps_out, True, False))
    else:
        layers.append(("Up-PS/Up-PS_CNN", cfg.cnn_size, cin, cfg.scale * cfg.scale * ps_out, True, False))
    cin = ps_out
    for i in range(cfg.reconstruct_layers - 1):
        layers.append(("R-CNN%d" % (i + 1), cfg.cnn_size, cin, cfg.reconstruct_filters, True, True))
        cin = cfg.reconstruct_filters
    layers.append(("R-CNN%d" % cfg.reconstruct_layers, cfg.cnn_size, cin, 1, False, False))
    return layers


def variable_names(cfg):
    """TF variable names a checkpoint of this config holds."""
    names = []
    for scope, k, cin, cout, bias, prelu in layer_table(cfg):
        base = scope.split("/")[-1]
        names.append(scope + "/conv_W")
        if bias:
            names.append(scope + "/conv_B")
        if cfg.depthwise_separable:
            names.append(scope + "/depthwise_W")
            names.append(scope + "/pointwise_W")
        if prelu:
            names.append("%s/prelu/%s_prelu" % (scope, base))
    return names


def he_init_weights(cfg, seed=0):
    """Random weights with the reference's default 'he' initialiser statistics
    (utilty.py:360-363: truncated normal, stddev sqrt(2/(k*k*cin))), bias 0,
    PReLU alpha 0.1 (tf_graph.py:91).  Small random biases/alphas are added so
    tests exercise those code paths."""
    g = np.random.RandomState(seed)
    w = {}
    for scope, k, cin, cout, bias, prelu in layer_table(cfg):
        base = scope.split("/")[-1]
        std = math.sqrt(2.0 / (k * k * cin))
        w[scope + "/conv_W"] = np.clip(g.randn(k, k, cin, cout), -2, 2).astype(np.float32) * np.float32(std)
        if cfg.depthwise_separable:
            w[scope + "/depthwise_W"] = (np.clip(g.randn(k, k, cin, 1), -2, 2) * math.sqrt(2.0 / (k * k * cin))
                                         ).astype(np.float32) + np.float32(1.0 / (k * k))
            w[scope + "/pointwise_W"] = (np.clip(g.randn(1, 1, cin, cout), -2, 2) * math.sqrt(2.0 / cin)
                                         ).astype(np.float32)
        if bias:
            w[scope + "/conv_B"] = (0.1 * g.randn(cout)).astype(np.float32)
        if prelu:
            w["%s/prelu/%s_prelu" % (scope, base)] = (0.1 + 0.05 * g.rand(cout)).astype(np.float32)
    return w


# ------------------------------------------------------------------ ops ----

def _t(a, dtype):
    return torch.from_numpy(np.array(a, copy=True, order="C")).to(dtype)


def conv2d_same(x_nchw, w_hwio, dtype):
    """tf.nn.conv2d(SAME, stride 1) with an HWIO filter (tf_graph.py:105)."""
    k = w_hwio.shape[0]
    w = _t(w_hwio, dtype).permute(3, 2, 0, 1).contiguous()  # OIHW; torch conv2d is cross-correlation too
    return F.conv2d(x_nchw, w, padding=k // 2)


def depthwise_same(x_nchw, w_hwi1, dtype):
    """Depthwise half of tf.nn.separable_conv2d, channel multiplier 1 (tf_graph.py:157-166)."""
    k, _, cin, _ = w_hwi1.shape
    w = _t(w_hwi1, dtype).permute(2, 3, 0, 1).contiguous()  # [cin,1,k,k]
    return F.conv2d(x_nchw, w, padding=k // 2, groups=cin)


def prelu(x_nchw, alpha, dtype):
    """relu(x) + alpha*(x-|x|)*0.5  (tf_graph.py:94)."""
    a = _t(alpha, dtype).view(1, -1, 1, 1)
    return torch.relu(x_nchw) + a * (x_nchw - torch.abs(x_nchw)) * 0.5


def depth_to_space(x_nchw, r):
    """tf.depth_to_space on NHWC == DCR: in_ch = (i*r + j)*C + c  (tf_graph.py:248)."""
    n, c, h, w = x_nchw.shape
    co = c // (r * r)
    x = x_nchw.view(n, r, r, co, h, w)           # [n, i, j, c, h, w]
    x = x.permute(0, 3, 4, 1, 5, 2).contiguous()  # [n, c, h, i, w, j]
    return x.view(n, co, h * r, w * r)


class Oracle:
    """Forward (and autograd backward) of the DCSCN graph on CPU."""

    def __init__(self, cfg, weights, dtype=torch.float32):
        self.cfg = cfg
        self.dtype = dtype
        self.w = weights
        self.table = layer_table(cfg)

    # one `build_conv` / `build_depthwise_separable_conv` (tf_graph.py:117-216)
    def _layer(self, scope, x, params=None, keep_prob=1.0, masks=None):
        cfg = self.cfg
        p = params if params is not None else self.w
        entry = [e for e in self.table if e[0] == scope][0]
        _, k, cin, cout, bias, use_prelu = entry
        base = scope.split("/")[-1]

        def get(name):
            v = p[name]
            return v if torch.is_tensor(v) else _t(v, self.dtype)

        if cfg.depthwise_separable:
            dw = get(scope + "/depthwise_W").permute(2, 3, 0, 1).contiguous()
            pw = get(scope + "/pointwise_W").permute(3, 2, 0, 1).contiguous()
            h = F.conv2d(x, dw, padding=k // 2, groups=cin)
            h = F.conv2d(h, pw)
        else:
            w = get(scope + "/conv_W").permute(3, 2, 0, 1).contiguous()
            h = F.conv2d(x, w, padding=k // 2)
        if bias:
            h = h + get(scope + "/conv_B").view(1, -1, 1, 1)
        if use_prelu:
            a = get("%s/prelu/%s_prelu" % (scope, base)).view(1, -1, 1, 1)
            h = torch.relu(h) + a * (h - torch.abs(h)) * 0.5
            if keep_prob < 1.0:
                # tf.nn.dropout: keep with prob keep_prob and scale by 1/keep_prob (tf_graph.py:130).
                # The mask is an explicit input so the CUDA path can be compared bit-for-bit on it.
                m = masks[scope]
                m = m if torch.is_tensor(m) else _t(m, self.dtype)
                h = h * m * (1.0 / keep_prob)
        return h

    def forward_nchw(self, x, x2, params=None, keep_prob=1.0, masks=None, return_intermediates=False):
        cfg = self.cfg
        inter = {}
        feats = []
        h = x
        for i in range(cfg.layers):
            h = self._layer("CNN%d" % (i + 1), h, params, keep_prob, masks)
            feats.append(h)
            inter["CNN%d" % (i + 1)] = h
        hc = torch.cat(feats, dim=1)                                   # DCSCN.py:259
        a1 = self._layer("A1", hc, params, keep_prob, masks)
        b1 = self._layer("B1", hc, params, keep_prob, masks)
        b2 = self._layer("B2", b1, params, keep_prob, masks)
        inter["A1"], inter["B1"], inter["B2"] = a1, b1, b2
        h = torch.cat([b2, a1], dim=1)                                 # DCSCN.py:281 ([H[-1], H[-3]])
        if cfg.scale == 4:                                             # DCSCN.py:298-304
            h = depth_to_space(self._layer("Up-PS/Up-PS_CNN", h, params), 2)
            inter["Up-PS"] = h
            h = depth_to_space(self._layer("Up-PS2/Up-PS2_CNN", h, params), 2)
            inter["Up-PS2"] = h
        else:
            h = depth_to_space(self._layer("Up-PS/Up-PS_CNN", h, params), cfg.scale)
            inter["Up-PS"] = h
        for i in range(cfg.reconstruct_layers - 1):
            h = self._layer("R-CNN%d" % (i + 1), h, params, keep_prob, masks)
        h = self._layer("R-CNN%d" % cfg.reconstruct_layers, h, params)
        inter["R-CNN"] = h
        y = h + x2                                                     # DCSCN.py:325
        if return_intermediates:
            return y, inter
        return y

    def forward(self, x_nhwc, x2_nhwc, return_intermediates=False):
        """x: [N,h,w,1], x2: [N,s*h,s*w,1] numpy -> y_ [N,s*h,s*w,1] numpy (same dtype family)."""
        with torch.no_grad():
            x = _t(x_nhwc, self.dtype).permute(0, 3, 1, 2).contiguous()
            x2 = _t(x2_nhwc, self.dtype).permute(0, 3, 1, 2).contiguous()
            out = self.forward_nchw(x, x2, return_intermediates=return_intermediates)
            if return_intermediates:
                y, inter = out
                return (y.permute(0, 2, 3, 1).contiguous().numpy(),
                        {k: v.permute(0, 2, 3, 1).contiguous().numpy() for k, v in inter.items()})
            return out.permute(0, 2, 3, 1).contiguous().numpy()

    # --------------------------------------------------- training step ----

    def trainable_names(self):
        """tf.trainable_variables() of the graph: W, B, alpha (and the DS filters)."""
        return variable_names(self.cfg)

    def l2_weight_names(self):
        """self.Weights: the conv_W of every layer (DCSCN.py:350; for DS graphs this is
        the dead conv_W, tf_graph.py:183,212)."""
        return [scope + "/conv_W" for scope, *_ in self.table]

    def loss_and_grads(self, x_nhwc, x2_nhwc, y_nhwc, keep_prob=1.0, masks=None, use_l1_loss=False):
        """mse, loss = image_loss + l2_decay*sum(sum(W^2)/2), d loss / d trainables (DCSCN.py:340-357,399);
        image_loss = mse, or mean|diff| with use_l1_loss (DCSCN.py:342-347)."""
        cfg = self.cfg
        params = {n: _t(self.w[n], self.dtype).clone().requires_grad_(True) for n in self.trainable_names()}
        x = _t(x_nhwc, self.dtype).permute(0, 3, 1, 2).contiguous()
        x2 = _t(x2_nhwc, self.dtype).permute(0, 3, 1, 2).contiguous()
        y = _t(y_nhwc, self.dtype).permute(0, 3, 1, 2).contiguous()
        y_ = self.forward_nchw(x, x2, params=params, keep_prob=keep_prob, masks=masks)
        diff = y_ - y
        mse = torch.mean(diff * diff)
        loss = torch.mean(torch.abs(diff)) if use_l1_loss else mse
        if cfg.l2_decay > 0:
            l2 = sum(torch.sum(params[n] * params[n]) / 2 for n in self.l2_weight_names())
            loss = loss + cfg.l2_decay * l2
        grads = torch.autograd.grad(loss, [params[n] for n in self.trainable_names()], allow_unused=True)
        g = {}
        for n, gr in zip(self.trainable_names(), grads):
            g[n] = (torch.zeros_like(params[n]) if gr is None else gr).detach().numpy()
        return float(mse.detach()), float(loss.detach()), g

    def clip_by_global_norm(self, grads):
        """tf.clip_by_global_norm(grads, clip_norm) (DCSCN.py:407): g * clip / max(norm, clip)."""
        clip = self.cfg.clipping_norm
        norm = math.sqrt(sum(float(np.sum(np.square(g.astype(np.float64)))) for g in grads.values()))
        if clip <= 0:
            return dict(grads), norm
        scale = clip / max(norm, clip)
        return {n: (g * g.dtype.type(scale)) for n, g in grads.items()}, norm

    def adam_step(self, grads, m, v, step, lr):
        """tf.train.AdamOptimizer update (DCSCN.py:388): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
        m = b1*m+(1-b1)*g; v = b2*v+(1-b2)*g*g; w -= lr_t*m/(sqrt(v)+eps).  Updates self.w in place."""
        cfg = self.cfg
        b1, b2, eps = cfg.beta1, cfg.beta2, cfg.epsilon
        lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        for n, g in grads.items():
            g = g.astype(np.float64)
            m[n] = b1 * m[n] + (1 - b1) * g
            v[n] = b2 * v[n] + (1 - b2) * g * g
            self.w[n] = (self.w[n].astype(np.float64) - lr_t * m[n] / (np.sqrt(v[n]) + eps)).astype(
                self.w[n].dtype)


# ------------------------------------------------- host pipeline restated ----
# (needed so the README PSNR table can pin the oracle end to end)

def convert_rgb_to_y(image):
    """utilty.py:142-149"""
    if len(image.shape) <= 2 or image.shape[2] == 1:
        return image
    xform = np.array([[65.738 / 256.0, 129.057 / 256.0, 25.064 / 256.0]])
    return image.dot(xform.T) + 16.0


def set_image_alignment(image, alignment):
    """utilty.py:196-208"""
    alignment = int(alignment)
    width, height = image.shape[1], image.shape[0]
    width = (width // alignment) * alignment
    height = (height // alignment) * alignment
    if image.shape[1] != width or image.shape[0] != height:
        image = image[:height, :width, :]
    if len(image.shape) >= 3 and image.shape[2] >= 4:
        image = image[:, :, 0:3]
    return image


def resize_image_by_pil(image, scale):
    """utilty.py:211-239, bicubic only.  Float arrays become PIL mode 'F' images."""
    from PIL import Image
    width, height = image.shape[1], image.shape[0]
    new_width = int(width * scale)
    new_height = int(height * scale)
    if len(image.shape) == 3 and image.shape[2] == 3:
        im = Image.fromarray(image, "RGB").resize([new_width, new_height], resample=Image.BICUBIC)
        return np.asarray(im)
    im = Image.fromarray(image.reshape(height, width))
    im = im.resize([new_width, new_height], resample=Image.BICUBIC)
    return np.asarray(im).reshape(new_height, new_width, 1)


def load_image(filename):
    """utilty.py:242-266 (imageio.imread replaced by PIL; same decoded pixels for PNG/BMP)."""
    from PIL import Image
    im = Image.open(filename)
    if im.mode not in ("L", "RGB", "RGBA"):
        im = im.convert("RGB")
    image = np.atleast_3d(np.asarray(im))
    if image.shape[2] >= 4:
        image = image[:, :, 0:3]
    return image


def flip(image, flip_type, invert=False):
    """utilty.py:595-617"""
    if flip_type == 0:
        return image
    if flip_type == 1:
        return np.flipud(image)
    if flip_type == 2:
        return np.fliplr(image)
    if flip_type == 3:
        return np.flipud(np.fliplr(image))
    if flip_type == 4:
        return np.rot90(image, 1 if invert is False else -1)
    if flip_type == 5:
        return np.rot90(image, -1 if invert is False else 1)
    if flip_type == 6:
        return np.flipud(np.rot90(image)) if invert is False else np.rot90(np.flipud(image), -1)
    if flip_type == 7:
        return np.flipud(np.rot90(image, -1)) if invert is False else np.rot90(np.flipud(image), 1)
    raise ValueError(flip_type)


def compute_psnr(image1, image2, border_size=0):
    """PSNR half of utilty.py:509-536 (rint, clip 0..255, shave border, skimage PSNR with
    data_range=255 == 10*log10(255^2 / mean((a-b)^2)) in float64)."""
    a = np.clip(np.rint(image1), 0, 255).astype(np.float32)
    b = np.clip(np.rint(image2), 0, 255).astype(np.float32)
    if a.ndim == 2:
        a = a.reshape(a.shape[0], a.shape[1], 1)
    if b.ndim == 2:
        b = b.reshape(b.shape[0], b.shape[1], 1)
    if border_size > 0:
        a = a[border_size:-border_size, border_size:-border_size, :]
        b = b[border_size:-border_size, border_size:-border_size, :]
    err = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    if err == 0:
        return float("inf")
    return 10.0 * math.log10(255.0 * 255.0 / err)


def do(oracle, input_image, bicubic_input_image, self_ensemble=8):
    """SuperResolution.do (DCSCN.py:547-586) with max_value 255."""
    s = oracle.cfg.scale
    h, w = input_image.shape[:2]
    if self_ensemble > 1:
        output = np.zeros([s * h, s * w, 1])
        for i in range(self_ensemble):
            image = flip(input_image, i)
            bic = flip(bicubic_input_image, i)
            y = oracle.forward(image.reshape(1, image.shape[0], image.shape[1], 1),
                               bic.reshape(1, s * image.shape[0], s * image.shape[1], 1))
            output += flip(y[0], i, invert=True)
        output /= self_ensemble
        return output
    return oracle.forward(input_image.reshape(1, h, w, 1), bicubic_input_image.reshape(1, s * h, s * w, 1))[0]


def build_inputs_for_evaluate(file_path, scale):
    """The host half of do_for_evaluate (DCSCN.py:672-696, loader.py:42-67):
    returns (input_y [h,w,1], bicubic_y [sh,sw,1], true_y [sh,sw,1])."""
    true_image = set_image_alignment(load_image(file_path), scale)
    if true_image.shape[2] == 3:
        input_y = resize_image_by_pil(convert_rgb_to_y(true_image), 1.0 / scale)
        true_y = convert_rgb_to_y(true_image)
    else:
        input_y = resize_image_by_pil(true_image, 1.0 / scale)
        true_y = true_image
    bicubic_y = resize_image_by_pil(input_y, scale)
    return input_y, bicubic_y, true_y


def do_for_evaluate(oracle, file_path, self_ensemble=8):
    """DCSCN.py:672-703 -> PSNR (border = scale, DCSCN.py:80-82)."""
    s = oracle.cfg.scale
    input_y, bicubic_y, true_y = build_inputs_for_evaluate(file_path, s)
    out = do(oracle, input_y, bicubic_y, self_ensemble)
    return compute_psnr(true_y, out, border_size=s)
