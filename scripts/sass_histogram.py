"""SASS opcode histogram of every kernel in libdcscn_b200.so (cuobjdump -sass), written to profiles/: the evidence that the
hot kernels issue tcgen05 (UTC*MMA), TMEM loads (LDTM), TMA (UTMALDG / UBLKCP) and mbarrier (SYNCS) instructions.

  python scripts/sass_histogram.py [out_file]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dcscn-super-resolution_b200", "csrc", "libdcscn_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "ELECT", "USETMAXREG",
       "UCGABAR", "MEMBAR", "ERRBAR", "HMMA", "FFMA", "FADD", "LDG", "STG", "LDS", "STS", "ATOMG", "RED", "DFMA", "DADD", "DMUL")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_sass_opcodes.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    parts = re.split(r"\n\s*Function : \S+", sass)[1:]
    lines = ["SASS opcode histogram of %s (cuobjdump -sass, sm_100a); full mnemonic incl. modifiers for the tcgen05 / TMA lines"
             % os.path.relpath(LIB, ROOT), ""]
    for nm, body in zip(name, parts):
        ops = collections.Counter()
        full = collections.Counter()
        for m in re.finditer(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body):
            op = m.group(1)
            ops[op.split(".")[0]] += 1
            if op.split(".")[0] in ("UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "UTCATOMSWS"):
                full[op] += 1
        total = sum(ops.values())
        short = nm.split("(")[0].replace("void ", "").replace("dcscn::", "")
        lines.append("%s   [%d instructions]" % (short, total))
        lines.append("   " + "  ".join("%s %d" % (k, ops[k]) for k in KEY if ops.get(k)))
        if full:
            lines.append("   " + "  ".join("%s x%d" % kv for kv in sorted(full.items())))
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
