"""Summarise an ncu report (read here, on the CPU box) into profiles/: a markdown table of the
metrics DESIGN.md cites and profiles/traffic.json (DRAM bytes per launch per kernel) that bench.py
reads for roofline.traffic.

usage: python tools/ncu_summary.py gpurun_out/<rep>.ncu-rep <workload>_<dtype> profiles/<name>.md"""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ lim regs"),
    ("launch__occupancy_limit_shared_mem", "occ lim smem"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def short(name):
    name = re.sub(r"^void ", "", name).replace("(int)", "")
    name = re.sub(r"\(.*$", "", name)
    return name.replace("scb::", "")


def main():
    rep, key, out_md = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    lines = ["| kernel | " + " | ".join(lbl for _, lbl in METRICS) + " |", "|---|" + "---|" * len(METRICS)]
    traffic = {}
    for r in data:
        name = short(r[col["Kernel Name"]])
        cells = []
        for m, _ in METRICS:
            if m not in col:
                cells.append("n/a")
                continue
            v, u = r[col[m]], units[col[m]]
            cells.append("%s %s" % (v, u) if u and u not in ("%",) else v)
        lines.append("| `%s` | " % name + " | ".join(cells) + " |")
        try:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_read.sum"]]]
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[col["dram__bytes_write.sum"]]]
            traffic[name.split("<")[0] if False else name] = rd + wr
        except Exception:
            pass
    with open(out_md, "w") as f:
        f.write("ncu --set full --clock-control none, report `%s` (per launch; cold-cache, serialised)\n\n" % os.path.basename(rep))
        f.write("\n".join(lines) + "\n")
    tpath = os.path.join(os.path.dirname(out_md), "traffic.json")
    allt = json.load(open(tpath)) if os.path.exists(tpath) else {}
    # bench.py looks kernels up by their base name
    base = {}
    for k, v in traffic.items():
        b = k.split("<")[0]
        if "k_lines" in k:
            b = "k_lines<-1>" if "-1>" in k else "k_lines<+1>"
        base[b] = v
    allt[key] = base
    json.dump(allt, open(tpath, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
