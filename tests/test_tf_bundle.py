"""TF V2 bundle reader / writer (helper/tf_bundle.py) against the known answers of the shipped checkpoints
 and a write -> read round trip.  CPU only."""
import os

import numpy as np
import pytest

from helper import tf_bundle

from conftest import GOLDEN

L12 = os.path.join(GOLDEN, "models", "dcscn_L12_F196to48_NIN_A64_PS_R1F32.ckpt")


def test_crc32c_known_vectors():
    # RFC 3720 test vectors
    assert tf_bundle.crc32c(b"") == 0
    assert tf_bundle.crc32c(b"123456789") == 0xE3069283
    assert tf_bundle.crc32c(bytes(32)) == 0x8A9136AA
    assert tf_bundle.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43


def test_l12_checkpoint_known_answers():
    r = tf_bundle.BundleReader(L12)
    assert len(r.entries) == 48  # 12 x (W,B,alpha) + A1,B1,B2 x 3 + Up-PS (W,B) + R-CNN1 W
    a = r.get_tensor("A1/conv_B", verify_crc=True)
    assert a.shape == (64,)
    assert np.float32(a[0]) == np.float32(4.2080207) and np.float32(a[1]) == np.float32(3.3013797)
    # masked crc32c of the raw tensor bytes recorded in the reference's own .index file
    assert r.entries["A1/conv_B"]["crc32c"] == 1450342233
    assert r.entries["CNN1/conv_W"]["crc32c"] == 4176728615 and r.shape("CNN1/conv_W") == [3, 3, 1, 196]
    assert r.entries["R-CNN1/conv_W"]["crc32c"] == 3061968345 and r.shape("R-CNN1/conv_W") == [3, 3, 96, 1]
    assert r.shape("Up-PS/Up-PS_CNN/conv_W") == [3, 3, 96, 384]
    for name in r.keys():
        r.get_tensor(name, verify_crc=True)


@pytest.mark.parametrize("name,count,params", [
    ("dcscn_L12_F196to48_Sc4_NIN_A64_PS_R1F32", 50, 2087102),
    ("dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32", 33, 27209),
    ("dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32", 61, None),
])
def test_other_checkpoints(name, count, params):
    r = tf_bundle.BundleReader(os.path.join(GOLDEN, "models", name + ".ckpt"))
    assert len(r.entries) == count
    if params is not None:
        assert sum(int(np.prod(r.shape(k))) for k in r.keys()) == params


def test_round_trip(tmp_path):
    g = np.random.RandomState(0)
    tensors = {"CNN%d/conv_W" % i: g.randn(3, 3, 5, 7).astype(np.float32) for i in range(1, 40)}
    tensors["scalar"] = np.float32(3.5).reshape(())
    tensors["A1/prelu/A1_prelu"] = g.rand(64).astype(np.float32)
    prefix = str(tmp_path / "m.ckpt")
    tf_bundle.write_bundle(prefix, tensors)
    r = tf_bundle.BundleReader(prefix)
    assert r.keys() == sorted(tensors.keys())
    for k, v in tensors.items():
        np.testing.assert_array_equal(r.get_tensor(k, verify_crc=True), v)


def test_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        tf_bundle.BundleReader(str(tmp_path / "missing.ckpt"))
    bad = tmp_path / "bad.ckpt.index"
    bad.write_bytes(b"\x00" * 100)
    with pytest.raises(ValueError):
        tf_bundle.BundleReader(str(tmp_path / "bad.ckpt"))
    # corrupt data -> crc mismatch
    prefix = str(tmp_path / "c.ckpt")
    tf_bundle.write_bundle(prefix, {"w": np.arange(8, dtype=np.float32)})
    p = prefix + ".data-00000-of-00001"
    raw = bytearray(open(p, "rb").read())
    raw[3] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        tf_bundle.BundleReader(prefix).get_tensor("w", verify_crc=True)


def test_scalar_tensors_round_trip_with_empty_shape(tmp_path):
    """beta1_power / beta2_power are rank-0 in the reference's checkpoints (shape proto without dims)."""
    from helper import tf_bundle
    prefix = str(tmp_path / "m.ckpt")
    tf_bundle.write_bundle(prefix, {"beta1_power": np.asarray(0.9 ** 5, dtype=np.float32), "w": np.ones((2, 3), np.float32)})
    r = tf_bundle.BundleReader(prefix)
    assert r.shape("beta1_power") == [] and r.get_tensor("beta1_power").shape == ()
    assert float(r.get_tensor("beta1_power")) == pytest.approx(0.9 ** 5, rel=1e-6)
    assert r.shape("w") == [2, 3]
