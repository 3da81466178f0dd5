"""Per-kernel SASS opcode histogram of libamphion_b200.so (evidence that the hot kernels are Blackwell-native:
UTCHMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk (TMA 1-D),
LDGSTS = cp.async, SYNCS = mbarrier, F*2 = FFMA2 + FMUL2 + FADD2, the packed fp32 pairs of add/mul/fma.f32x2).  Runs on the CPU box: python scripts/sass_histogram.py > profiles/r2_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "amphion_b200", "libamphion_b200.so")
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "HMMA", "FFMA", "F*2", "MUFU", "LDG", "STG", "LDS", "STS"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    name, hist = None, collections.OrderedDict()
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::", "", name)
            name = re.sub(r"\(.*", "", name)
            hist[name] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m and name:
            hist[name][m.group(1)] += 1
            if m.group(1) in ("FFMA2", "FMUL2", "FADD2"):
                hist[name]["F*2"] += 1
    print("%-64s %7s  %s" % ("kernel", "instrs", "  ".join("%s" % k for k in KEYS)))
    for name, h in hist.items():
        total = sum(v for k, v in h.items() if k != "F*2")
        print("%-64s %7d  %s" % (name[:64], total, "  ".join("%*d" % (len(k), h.get(k, 0)) for k in KEYS)))


if __name__ == "__main__":
    sys.exit(main())
