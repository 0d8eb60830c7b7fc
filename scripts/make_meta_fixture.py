"""Extracts the FORWARD sub-graph of the MetaGraphDefs the reference ships (`/root/reference/models/*.ckpt.meta`, written by
the reference's own `tf.train.Saver`) into small JSON fixtures under tests/golden/meta/.

The .meta files are serialized protobufs (MetaGraphDef -> GraphDef -> NodeDef); TensorFlow is not installed, so the wire
format is decoded directly (varint / length-delimited fields; field numbers from tensorflow/core/framework/*.proto).
Only nodes the network output depends on are kept (placeholders, variables, Conv2D, DepthwiseConv2dNative, Add, PReLU's
Relu/Abs/Sub/Mul, the dropout sub-graph, ConcatV2, DepthToSpace), with the attributes that decide the arithmetic.
tests/test_oracle_meta_graph.py executes these graphs op by op and holds the oracle to them.

  python scripts/make_meta_fixture.py            # rewrites tests/golden/meta/*.json (needs /root/reference)
"""
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODELS = "/root/reference/models"
OUT = os.path.join(ROOT, "tests", "golden", "meta")
MODELS = ["dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32", "dcscn_L7_F32to8_G1.20_Sc3_NIN_A24_B8_PS_R1F32",
          "dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_R1F32", "dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32"]


def varint(b, i):
    r = s = 0
    while True:
        c = b[i]
        i += 1
        r |= (c & 0x7F) << s
        s += 7
        if not c & 0x80:
            return r, i


def fields(b):
    """[(field number, wire type, value)] of one serialized message."""
    i, out = 0, []
    while i < len(b):
        k, i = varint(b, i)
        f, wt = k >> 3, k & 7
        if wt == 0:
            v, i = varint(b, i)
        elif wt == 1:
            v, i = b[i:i + 8], i + 8
        elif wt == 2:
            n, i = varint(b, i)
            v, i = b[i:i + n], i + n
        elif wt == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError("wire type %d" % wt)
        out.append((f, wt, v))
    return out


def packed_varints(b):
    i, out = 0, []
    while i < len(b):
        v, i = varint(b, i)
        out.append(v)
    return out


def shape_dims(b):
    """TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}."""
    dims = []
    for f, _, v in fields(b):
        if f == 2:
            size = 0
            for f2, _, v2 in fields(v):
                if f2 == 1:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return dims


def tensor_value(b):
    """TensorProto -> {"dtype": enum, "shape": [...], "values": [...]} (float_val = 5, int_val = 7, tensor_content = 4)."""
    dtype, shape, vals = 0, [], []
    for f, wt, v in fields(b):
        if f == 1:
            dtype = v
        elif f == 2:
            shape = shape_dims(v)
        elif f == 4:
            fmt = {1: "f", 3: "i", 9: "q"}.get(dtype)
            if fmt is None:
                raise ValueError("tensor_content of dtype %d" % dtype)
            vals = list(struct.unpack("<%d%s" % (len(v) // struct.calcsize(fmt), fmt), v))
        elif f == 5:
            vals += list(struct.unpack("<%df" % (len(v) // 4), v)) if wt == 2 else [struct.unpack("<f", v)[0]]
        elif f == 7:
            vals += packed_varints(v) if wt == 2 else [v]
    return {"dtype": dtype, "shape": shape, "values": vals}


def attr_value(b):
    """AttrValue oneof: list = 1, s = 2, i = 3, f = 4, b = 5, type = 6, shape = 7, tensor = 8."""
    for f, wt, v in fields(b):
        if f == 2:
            return v.decode("utf-8", "replace")
        if f == 3:
            return v
        if f == 4:
            return struct.unpack("<f", v)[0]
        if f == 5:
            return bool(v)
        if f == 6:
            return {"type": v}
        if f == 7:
            return shape_dims(v)
        if f == 8:
            return tensor_value(v)
        if f == 1:  # ListValue: s = 2, i = 3 (packed), f = 4, shape = 7
            out = []
            for f2, wt2, v2 in fields(v):
                if f2 == 3:
                    out += packed_varints(v2) if wt2 == 2 else [v2]
                elif f2 == 2:
                    out.append(v2.decode("utf-8", "replace"))
            return out
    return None


KEEP_ATTRS = ("padding", "strides", "data_format", "dilations", "block_size", "N", "shape", "value", "DstT", "SrcT")


def graph_nodes(meta_bytes):
    graph_def = [v for f, _, v in fields(meta_bytes) if f == 2][0]       # MetaGraphDef.graph_def
    producer = None
    nodes = []
    for f, _, v in fields(graph_def):
        if f == 4:                                                        # GraphDef.versions {producer = 1}
            producer = dict((a, c) for a, _, c in fields(v)).get(1)
        if f != 1:                                                        # GraphDef.node
            continue
        n = {"name": "", "op": "", "input": [], "attr": {}}
        for f2, _, v2 in fields(v):
            if f2 == 1:
                n["name"] = v2.decode()
            elif f2 == 2:
                n["op"] = v2.decode()
            elif f2 == 3:
                n["input"].append(v2.decode())
            elif f2 == 5:                                                 # map<string, AttrValue> entry
                kv = dict((a, c) for a, _, c in fields(v2))
                key = kv[1].decode()
                if key in KEEP_ATTRS:
                    n["attr"][key] = attr_value(kv.get(2, b""))
        nodes.append(n)
    return nodes, producer


def forward_subgraph(nodes):
    by = {n["name"]: n for n in nodes}
    if "output" in by:                    # DCSCN.py:325 names the residual add "output"; older graphs call it "add"
        root = "output"
    else:
        root = [n["name"] for n in nodes if n["op"] == "Add" and "x2" in n["input"] and not n["name"].startswith("gradients")][0]
    seen, stack = [], [root]
    mark = set()
    while stack:
        k = stack.pop().lstrip("^").split(":")[0]
        if k in mark:
            continue
        mark.add(k)
        seen.append(k)
        stack.extend(by[k]["input"])
    keep = [n for n in nodes if n["name"] in mark]                        # file order = construction order
    return root, keep


def main():
    os.makedirs(OUT, exist_ok=True)
    for m in MODELS:
        raw = open(os.path.join(REF_MODELS, m + ".ckpt.meta"), "rb").read()
        nodes, producer = graph_nodes(raw)
        root, keep = forward_subgraph(nodes)
        doc = {"source": "models/%s.ckpt.meta of the reference (MetaGraphDef written by tf.train.Saver, GraphDef producer %s)" % (m, producer),
               "generator": "scripts/make_meta_fixture.py", "root": root, "nodes": keep}
        path = os.path.join(OUT, m + ".json")
        with open(path, "w") as f:
            json.dump(doc, f, separators=(",", ":"))
        print("%s: %d of %d nodes, root %s, %d bytes" % (m, len(keep), len(nodes), root, os.path.getsize(path)))


if __name__ == "__main__":
    sys.exit(main())
