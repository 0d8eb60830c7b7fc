"""A tiny interpreter for the forward sub-graphs of the reference's shipped MetaGraphDefs (tests/golden/meta/*.json, made by
scripts/make_meta_fixture.py from /root/reference/models/*.ckpt.meta).  TEST INFRASTRUCTURE.

Each TF op type that occurs is implemented in plain numpy (float64) from TensorFlow's documented semantics, independent of
oracle/dcscn_oracle.py (no torch, no shared helper), so "oracle == this interpreter on the reference's own serialized graph"
pins the oracle's op order, operand wiring, padding, concat order, depth_to_space layout and PReLU form to bytes the
reference ships rather than to a reading of its Python source."""
import json

import numpy as np


def _conv2d_nhwc(x, w, padding):
    """tf.nn.conv2d, stride 1, NHWC input, HWIO filter, cross-correlation; SAME pads (k-1)//2 before and k//2 after."""
    kh, kw, cin, cout = w.shape
    assert x.shape[3] == cin
    if padding == "SAME":
        x = np.pad(x, ((0, 0), ((kh - 1) // 2, kh // 2), ((kw - 1) // 2, kw // 2), (0, 0)))
    else:
        assert padding == "VALID"
    n, hp, wp, _ = x.shape
    oh, ow = hp - kh + 1, wp - kw + 1
    out = np.zeros((n, oh, ow, cout), dtype=np.float64)
    for dy in range(kh):
        for dx in range(kw):
            out += x[:, dy:dy + oh, dx:dx + ow, :] @ w[dy, dx]
    return out


def _depthwise_nhwc(x, w, padding):
    """tf.nn.depthwise_conv2d_native: filter [kh, kw, cin, multiplier], output channel = c * multiplier + m."""
    kh, kw, cin, mult = w.shape
    assert x.shape[3] == cin
    if padding == "SAME":
        x = np.pad(x, ((0, 0), ((kh - 1) // 2, kh // 2), ((kw - 1) // 2, kw // 2), (0, 0)))
    n, hp, wp, _ = x.shape
    oh, ow = hp - kh + 1, wp - kw + 1
    out = np.zeros((n, oh, ow, cin, mult), dtype=np.float64)
    for dy in range(kh):
        for dx in range(kw):
            out += x[:, dy:dy + oh, dx:dx + ow, :, None] * w[dy, dx][None, None, None, :, :]
    return out.reshape(n, oh, ow, cin * mult)


def _depth_to_space_nhwc(x, b):
    """tf.depth_to_space (NHWC): out[n, h*b + i, w*b + j, c] = in[n, h, w, (i*b + j)*C + c]."""
    n, h, w, d = x.shape
    c = d // (b * b)
    return x.reshape(n, h, w, b, b, c).transpose(0, 1, 3, 2, 4, 5).reshape(n, h * b, w * b, c)


class GraphInterpreter:
    def __init__(self, fixture_path):
        doc = json.load(open(fixture_path))
        self.root = doc["root"]
        self.nodes = {n["name"]: n for n in doc["nodes"]}
        self.order = [n["name"] for n in doc["nodes"]]

    def variables(self):
        return {k: tuple(n["attr"]["shape"]) for k, n in self.nodes.items() if n["op"] == "VariableV2"}

    def run(self, feeds, weights, fetch=None):
        """feeds: {placeholder: array}; weights: {variable name: array}; returns {node name: value} for `fetch` names
        (default: the root only)."""
        memo = {}

        def ev(name):
            name = name.lstrip("^").split(":")[0]
            if name in memo:
                return memo[name]
            n = self.nodes[name]
            op, a = n["op"], n["attr"]
            i = [ev(k) for k in n["input"]]
            if op == "Placeholder":
                v = np.asarray(feeds[name], dtype=np.float64)
            elif op == "VariableV2":
                v = np.asarray(weights[name], dtype=np.float64)
                assert tuple(v.shape) == tuple(a["shape"]), (name, v.shape, a["shape"])
            elif op == "Identity":
                v = i[0]
            elif op == "Const":
                t = a["value"]
                v = np.asarray(t["values"], dtype=np.float64 if t["dtype"] == 1 else np.int64)
                v = v.reshape(t["shape"]) if t["shape"] else v.reshape(())
            elif op == "Conv2D":
                assert a["strides"] == [1, 1, 1, 1] and a["data_format"] == "NHWC" and a.get("dilations", [1, 1, 1, 1]) == [1, 1, 1, 1]
                v = _conv2d_nhwc(i[0], i[1], a["padding"])
            elif op == "DepthwiseConv2dNative":
                assert a["strides"] == [1, 1, 1, 1] and a["data_format"] == "NHWC"
                v = _depthwise_nhwc(i[0], i[1], a["padding"])
            elif op in ("Add", "AddV2"):
                v = i[0] + i[1]
            elif op == "Sub":
                v = i[0] - i[1]
            elif op == "Mul":
                v = i[0] * i[1]
            elif op == "RealDiv":
                v = i[0] / i[1]
            elif op == "Relu":
                v = np.maximum(i[0], 0.0)
            elif op == "Abs":
                v = np.abs(i[0])
            elif op == "Floor":
                v = np.floor(i[0])
            elif op == "Shape":
                v = np.asarray(i[0].shape, dtype=np.int64)
            elif op == "RandomUniform":
                v = np.zeros(tuple(int(d) for d in i[0]), dtype=np.float64)   # u = 0: with keep = 1 every element is kept
            elif op == "GreaterEqual":
                v = i[0] >= i[1]
            elif op == "Cast":
                v = i[0].astype(np.float64)
            elif op == "ConcatV2":
                assert a["N"] == len(i) - 1
                v = np.concatenate(i[:-1], axis=int(i[-1]))
            elif op == "DepthToSpace":
                assert a["data_format"] == "NHWC"
                v = _depth_to_space_nhwc(i[0], a["block_size"])
            else:
                raise NotImplementedError("op %s (%s)" % (op, name))
            memo[name] = v
            return v

        names = [self.root] if fetch is None else list(fetch)
        return {k: ev(k) for k in names}
