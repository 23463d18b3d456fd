#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel JSON summaries kept under profiles/,
and (with --traffic) refresh profiles/traffic.json, which bench.py reads for `roofline.traffic`.

    python scripts/ncu_summary.py gpurun_out/r1d_full_raw.csv profiles/r1d_rollout_ncu_summary.json \
        --work "2368 lanes x 1024 cells x 64 steps" --cell-steps 155189248 --traffic
"""
import argparse
import csv
import json
import os
import re

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv"); ap.add_argument("out_json")
    ap.add_argument("--work", default=""); ap.add_argument("--command", default="")
    ap.add_argument("--cell-steps", type=float, default=0.0, help="cell-steps (or vehicle-steps) per ARZ launch")
    ap.add_argument("--traffic", action="store_true")
    ap.add_argument("--suffix", default="", help="appended to the traffic.json keys, e.g. _k1 for the store-every-state mode")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.raw_csv)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    out = []
    for r in rows[2:]:
        k = {"Kernel Name": r[ix["Kernel Name"]]}
        for m in KEEP:
            if m in ix:
                v, u = float(r[ix[m]].replace(",", "")), units[ix[m]]
                if u in SCALE:
                    v, u = v * SCALE[u], "byte"
                k[m] = v
                if u:
                    k.setdefault("units", {})[m] = u
        st = {re.sub(r"smsp__average_warps_issue_stalled_|_per_issue_active.ratio", "", h): float(r[ix[h]]) for h in stall}
        k["stalls_per_issue"] = {n: round(v, 3) for n, v in sorted(st.items(), key=lambda x: -x[1]) if v >= 0.05}
        if a.cell_steps and "arz" in k["Kernel Name"]:
            k["thread_inst_per_cell_step"] = k["smsp__inst_executed.sum"] * 32 / a.cell_steps
            k["dram_bytes_per_cell_step"] = (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) / a.cell_steps
        out.append(k)
    json.dump({"command": a.command, "work": a.work, "kernels": out}, open(a.out_json, "w"), indent=1)
    if a.traffic:
        path = os.path.join(os.path.dirname(a.out_json), "traffic.json")
        t = json.load(open(path)) if os.path.exists(path) else {}
        for k in out:
            name = re.match(r"void (\w+)<(\w+)", k["Kernel Name"])
            if not name:
                continue
            key = name.group(1).replace("_reg_kernel", "").replace("_kernel", "") + ("_f64" if name.group(2) == "double" else "_f32") + a.suffix
            t[key] = {"dram_bytes_per_launch": k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"], "work": a.work,
                      "source": os.path.basename(a.out_json)}
            if a.cell_steps and "arz" in key:
                t[key]["dram_bytes_per_cell_step"] = t[key]["dram_bytes_per_launch"] / a.cell_steps
        json.dump(t, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()
