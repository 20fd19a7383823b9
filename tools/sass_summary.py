"""Counts of the Blackwell-native SASS mnemonics per kernel of libmsda_b200.so (cuobjdump -sass): UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores, UBLKRED = cp.reduce.async.bulk, REDG.*x4 = 128-bit
vector reductions.  Runs without a GPU.  Writes profiles/sass_summary.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "ziragroundingdino_b200", "_lib", "libmsda_b200.so")
PAT = [("UTC*MMA", r"\bUTC[A-Z]*MMA\b"), ("LDTM", r"\bLDTM\b"), ("STTM", r"\bSTTM\b"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"),
       ("UBLKRED", r"\bUBLKRED"), ("REDG.F32x4", r"\bREDG\.E\.ADD\.F32x4"), ("REDG.other", r"\bREDG\.E\.ADD\.(?!F32x4)"),
       ("HMMA(legacy)", r"\bHMMA\b"), ("LDG.128", r"\bLDG\.E\.128|LDG\.E\.(?:[A-Z]+\.)*128"), ("SHFL", r"\bSHFL\b")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        for name, pat in PAT:
            if re.search(pat, line):
                cur[name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    out = ["SASS mnemonic counts per kernel of ziragroundingdino_b200/_lib/libmsda_b200.so (sm_100a), from `cuobjdump -sass`",
           "%-9s %5s %5s %8s %8s %8s %11s %10s %6s  kernel" % tuple(n for n, _ in PAT[:9])]
    tot = collections.Counter()
    for (mangled, c), name in zip(kernels.items(), demangle):
        tot.update(c)
        if not any(c[n] for n, _ in PAT[:9]):
            continue
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))[:110]
        out.append("%-9d %5d %5d %8d %8d %8d %11d %10d %6d  %s" % (*(c[n] for n, _ in PAT[:9]), short))
    out.append("%-9d %5d %5d %8d %8d %8d %11d %10d %6d  TOTAL over %d kernels" % (*(tot[n] for n, _ in PAT[:9]), len(kernels)))
    text = "\n".join(out) + "\n"
    path = os.path.join(ROOT, "profiles", "sass_summary.txt")
    open(path, "w").write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
