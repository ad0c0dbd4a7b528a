"""Opcode histogram per kernel of the built library (static SASS instruction counts): the evidence that the hot kernels use
tcgen05 (UTCHMMA / UTCBAR / LDTM / STTM), bulk copies (UBLKCP), mbarriers (SYNCS) and packed f32x2 (FFMA2 / FMUL2 / FADD2).
Usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt   (needs cuobjdump and c++filt on PATH; no GPU)"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "vi-hmc_b200", "vihmc", "libvihmc.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FFMA2", "FMUL2", "FADD2", "FFMA", "MUFU",
        "LDS", "STS", "LDG", "STG", "HMMA", "IMAD", "SHFL", "BAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["__total"] += 1
    names = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print("# cuobjdump -sass vi-hmc_b200/vihmc/libvihmc.so: opcode histogram per kernel (static instruction counts), tools/sass_summary.py")
    print("# UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk, SYNCS = mbarrier ops,")
    print("# FFMA2/FMUL2/FADD2 = packed f32x2")
    print("kernel | total | " + " | ".join(COLS))
    for (mangled, c), name in zip(kernels.items(), names):
        short = re.sub(r"\(.*", "", name)
        print(f"{short} | {c['__total']} | " + " | ".join(str(c[k]) for k in COLS))


if __name__ == "__main__":
    sys.exit(main())
