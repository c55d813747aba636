"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): time and launches per kernel, own vs stock.
Usage: python profiles/launch_summary.py launches.csv [skip_first_n_launches | first_kernel_of_a_step] > summary.txt"""
import csv, re, sys
from collections import defaultdict

import os, subprocess
LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "rgb-d-instance-segmentation_b200", "csrc", "librgbd_b200.so")
# kernel names of THIS library (demangled, without namespace / template arguments), from its own ELF
_sym = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
OWN_NAMES = set(re.findall(r"_cu_[0-9a-f]{8}\d+([a-z][a-z0-9_]*?_kernel)(?:I|E|v)", _sym))
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
# second argument: launches to skip, or a kernel name -- then the summary starts at that kernel's LAST launch (= last step)
skip = 0
if len(sys.argv) > 2:
    if sys.argv[2].isdigit():
        skip = int(sys.argv[2])
    else:
        skip = max(i for i, r in enumerate(rows[1:]) if sys.argv[2] in r[ki])
t, n = defaultdict(float), defaultdict(int)
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki])
    name = re.sub(r"^void ", "", name)
    m = re.match(r"<unnamed>::([a-z0-9_]+)", name)
    own = bool(m) and m.group(1) in OWN_NAMES
    key = ("own   " if own else "stock ") + name[:110]
    t[key] += float(r[vi].replace(",", "")) / 1e3
    n[key] += 1
tot = sum(t.values())
own_t = sum(v for k, v in t.items() if k.startswith("own"))
print(f"launches {sum(n.values())}, total {tot / 1e3:.2f} ms (cold-cache, serialised), own kernels {own_t / 1e3:.2f} ms = {100 * own_t / tot:.1f} %")
for k, v in sorted(t.items(), key=lambda kv: -kv[1])[:60]:
    print(f"{v / 1e3:9.3f} ms {100 * v / tot:5.1f} % x{n[k]:<5d} {k}")
