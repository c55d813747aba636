"""Turn an `ncu --set full` capture into the committed summaries bench.py and the judge read.

    python profiles/ncu_extract.py gpurun_out/r02_full.ncu-rep 32 "<how the capture was taken>"

writes profiles/r02_ncu_full_b<batch>.txt (selected raw metrics per kernel launch) and profiles/r02_ncu.json
({commit, batch, source, tc_util_pct{kernel: [..]}, dram_bytes_per_launch{kernel: ..}}), which bench.py loads for its
static `tc_util_pct_ncu` and `roofline.traffic` fields (never measured inside a bench run)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_dim_x"]


def short(name: str) -> str:
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|^void ", "", name)
    name = re.sub(r"\(.*", "", name)
    return re.sub(r"<[^>]*>$", "", name)          # template arguments: conv3x3_2cta_kernel<0> -> conv3x3_2cta_kernel


def to_bytes(value: str, unit: str) -> float:
    v = float(value.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    rep, batch, how = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    txt = [f"ncu --set full --clock-control none --import-source on; {how}; repo at {commit}", ""]
    tc, dram = {}, {}
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        k = short(r[col["Kernel Name"]])
        txt.append(f"  Kernel Name = {k}")
        for m in KEEP:
            if m in col:
                txt.append(f"  {m} [{units[col[m]]}] = {r[col[m]]}")
        txt.append("")
        m = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
        if m in col and r[col[m]] not in ("", "n/a"):
            tc.setdefault(k, []).append(round(float(r[col[m]].replace(",", "")), 1))
        if "dram__bytes_read.sum" in col:
            b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            dram.setdefault(k, []).append(b)
    out_txt = os.path.join(ROOT, "profiles", f"r02_ncu_full_b{batch}.txt")
    open(out_txt, "w").write("\n".join(txt) + "\n")
    js = {"commit": commit, "batch": batch, "source": os.path.relpath(out_txt, ROOT), "how": how,
          "tc_util_pct": tc, "dram_bytes_per_launch": {k: sum(v) / len(v) for k, v in dram.items()}}
    json.dump(js, open(os.path.join(ROOT, "profiles", "r02_ncu.json"), "w"), indent=1)
    print(json.dumps(js, indent=1))


if __name__ == "__main__":
    main()
