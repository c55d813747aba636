"""Per-kernel SASS instruction counts of csrc/librgbd_b200.so (tcgen05 / TMEM / TMA evidence).

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt

Counts the Blackwell-specific mnemonics per kernel from `cuobjdump -sass`: UTCHMMA (tcgen05.mma, incl. .2CTA),
UTCBAR (tcgen05.commit), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA tensor load / store), UTCCP, SYNCS
(mbarrier), plus the classic HMMA (mma.sync: must be 0) for contrast."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "rgb-d-instance-segmentation_b200", "csrc", "librgbd_b200.so")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCCP", "SYNCS", "HMMA", "LDSM", "MUFU", "ATOMS", "RED", "ATOMG"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.OrderedDict()
    cur = None
    n_ins = collections.Counter()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*", "", cur)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            n_ins[cur] += 1
            op = m.group(1)
            for mn in MNEMONICS:
                if op.startswith(mn):
                    counts[cur][mn + (".2CTA" if ".2CTA" in op and mn == "UTCHMMA" else "")] += 1
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "LDSM", "MUFU"]
    print(f"cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (sm_100a; one row per kernel; instruction counts)")
    print(f"{'kernel':58s} {'instr':>6s} " + " ".join(f"{c:>12s}" for c in cols))
    tot = collections.Counter()
    for k, c in counts.items():
        print(f"{k[:58]:58s} {n_ins[k]:6d} " + " ".join(f"{c.get(col, 0):12d}" for col in cols))
        tot.update(c)
    print(f"{'TOTAL':58s} {sum(n_ins.values()):6d} " + " ".join(f"{tot.get(col, 0):12d}" for col in cols))


if __name__ == "__main__":
    sys.exit(main())
