"""
Builds the checkpoint fixtures under tests/golden/models/ from the reference's shipped
`models/*.ckpt` bundles (run once in the build container, where /root/reference exists):
the trainable variables only (Adam slots and beta powers dropped, 3x smaller), re-written
with helper/tf_bundle.write_bundle in the same TF V2 bundle format and verified bit-exact.
Also copies the Set5 / Set14 evaluation images (assets, not sources) to tests/golden/data/.
"""
import os
import shutil
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dcscn-super-resolution_b200"))
from helper import tf_bundle  # noqa: E402

REF = "/root/reference"
MODELS = [
    "dcscn_L12_F196to48_NIN_A64_PS_R1F32",
    "dcscn_L12_F196to48_Sc4_NIN_A64_PS_R1F32",
    "dcscn_L7_F32to8_G1.20_NIN_A24_B8_PS_R1F32",
    "dcscn_L7_F32to8_G1.20_Sc3_NIN_A24_B8_PS_R1F32",
    "dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_R1F32",
    "dcscn_L7_F32to8_G1.20_Sc4_NIN_A24_B8_PS_DS_R1F32",
    # round 2: the remaining shipped variants (README.md:80,100,132 use --layers=8 --filters=96)
    "dcscn_L12_F196to48_Sc3_NIN_A64_PS_R1F32",
    "dcscn_L8_F96to48_NIN_A64_PS_R1F32",
    "dcscn_L8_F96to48_Sc3_NIN_A64_PS_R1F32",
    "dcscn_L8_F96to48_Sc4_NIN_A64_PS_R1F32",
]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden", "models")
    os.makedirs(out_dir, exist_ok=True)
    for m in MODELS:
        r = tf_bundle.BundleReader(os.path.join(REF, "models", m + ".ckpt"))
        keep = {k: r.get_tensor(k, verify_crc=True) for k in r.keys()
                if "/Adam" not in k and k not in ("beta1_power", "beta2_power")}
        prefix = os.path.join(out_dir, m + ".ckpt")
        tf_bundle.write_bundle(prefix, keep)
        r2 = tf_bundle.BundleReader(prefix)
        for k, v in keep.items():
            assert np.array_equal(r2.get_tensor(k, verify_crc=True), v), k
        print("%s: %d variables, %d params" % (m, len(keep), sum(v.size for v in keep.values())))
    for ds in ("set5", "set14"):
        dst = os.path.join(ROOT, "tests", "golden", "data", ds)
        os.makedirs(dst, exist_ok=True)
        for f in sorted(os.listdir(os.path.join(REF, "data", ds))):
            if not f.startswith("."):
                shutil.copyfile(os.path.join(REF, "data", ds, f), os.path.join(dst, f))
        print(ds, len(os.listdir(dst)), "images")


if __name__ == "__main__":
    main()
