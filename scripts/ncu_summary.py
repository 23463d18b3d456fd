#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` into the per-kernel JSON summaries kept under profiles/,
and (with --mix) refresh profiles/kernel_mix.json, which bench.py reads for `roofline.traffic` (DRAM bytes per
update) and `roofline_fp64` (fp64-pipe / total thread-instructions per update).

    python scripts/ncu_summary.py gpurun_out/r2a_full_raw.csv profiles/r2a_rollout_ncu_summary.json \
        --work "6560 lanes x 1024 cells x 256 steps" --cell-steps 1719664640 --vehicle-steps 268435456 --mix --suffix _k1
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
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_elapsed.avg", "sm__cycles_active.avg",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed"]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw_csv"); ap.add_argument("out_json")
    ap.add_argument("--work", default=""); ap.add_argument("--command", default="")
    ap.add_argument("--cell-steps", type=float, default=0.0, help="cell-steps (or vehicle-steps) per ARZ launch")
    ap.add_argument("--vehicle-steps", type=float, default=0.0, help="vehicle-steps per IDM launch")
    ap.add_argument("--traffic", action="store_true")
    ap.add_argument("--mix", action="store_true", help="refresh profiles/kernel_mix.json")
    ap.add_argument("--idm-suffix", default="", help="kernel_mix.json key suffix of the IDM kernels, e.g. _k16")
    ap.add_argument("--sms", type=int, default=148)
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
        units_ = a.cell_steps if "arz" in k["Kernel Name"] else (a.vehicle_steps if "idm" in k["Kernel Name"] else 0.0)
        if units_:
            cyc = k.get("smsp__cycles_elapsed.avg", 0.0)
            ar = sum(k.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % o, 0.0) for o in ("dadd", "dfma", "dmul"))
            k["thread_inst_per_update"] = k["smsp__inst_executed.sum"] * k.get("smsp__thread_inst_executed_per_inst_executed.ratio", 32.0) / units_
            k["dram_bytes_per_update"] = (k["dram__bytes_read.sum"] + k["dram__bytes_write.sum"]) / units_
            k["fp64_arith_inst_per_update"] = ar * cyc / units_                      # DADD + DFMA + DMUL thread-instructions
            # everything the fp64 pipe executed (arithmetic + DSETP / conversions): pipe-active cycles x 16 lanes per SMSP x 4
            k["fp64_pipe_inst_per_update"] = (k.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", 0.0) / 100.0
                                              * k.get("sm__cycles_active.avg", 0.0) * 64.0 * a.sms / units_)
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
    if a.mix:
        path = os.path.join(os.path.dirname(a.out_json), "kernel_mix.json")
        t = json.load(open(path)) if os.path.exists(path) else {}
        for k in out:
            name = re.match(r"void (\w+)<(\w+)", k["Kernel Name"])
            if not name or "thread_inst_per_update" not in k:
                continue
            kern = name.group(1).replace("_reg_kernel", "").replace("_kernel", "")
            key = kern + ("_f64" if name.group(2) == "double" else "_f32") + (a.suffix if "arz" in kern else a.idm_suffix)
            t[key] = {m: k[m] for m in ("thread_inst_per_update", "dram_bytes_per_update", "fp64_arith_inst_per_update",
                                        "fp64_pipe_inst_per_update")}
            t[key].update(work=a.work, source=os.path.basename(a.out_json), registers=k.get("launch__registers_per_thread"),
                          fp64_pipe_pct=k.get("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                          issue_pct=k.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          dram_pct=k.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"))
        json.dump(t, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1)[:3000])


if __name__ == "__main__":
    main()
