"""Turn ncu output into the small, committed summaries under profiles/ (harness, not product code).

  python tools/ncu_summarize.py launches <ncu --csv log> <out.jsonl> <out_summary.json> ["command line for the header"]
      parses the CSV launch list of  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
  python tools/ncu_summarize.py full <raw.csv from `ncu -i rep --page raw --csv`> <out.csv> ["header"]
      keeps the metrics that matter for an HBM-bound streaming kernel, one column per profiled launch
"""
import csv
import io
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "launch__occupancy_limit_blocks",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second",
]


def csv_rows(path):
    text = open(path, errors="replace").read()
    start = text.find('"ID"')
    if start < 0:
        raise SystemExit(f"{path}: no ncu CSV table found")
    return list(csv.reader(io.StringIO(text[start:])))


def launches(log, out_jsonl, out_summary, header=""):
    rows = csv_rows(log)
    head = rows[0]
    col = {name: i for i, name in enumerate(head)}
    per = {}
    for r in rows[1:]:
        if len(r) < len(head) or not r[col["ID"]].isdigit():
            continue
        k = int(r[col["ID"]])
        d = per.setdefault(k, {"id": k, "kernel": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        name, unit, val = r[col["Metric Name"]], r[col["Metric Unit"]], float(r[col["Metric Value"]].replace(",", ""))
        if name == "gpu__time_duration.sum":
            d["time_us"] = val * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        elif name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1.0)
            d["dram_read_bytes" if "read" in name else "dram_write_bytes"] = val * scale
    items = [per[k] for k in sorted(per)]
    with open(out_jsonl, "w") as f:
        if header:
            f.write("# " + header + "\n")
        for d in items:
            f.write(json.dumps(d) + "\n")
    total = sum(d.get("time_us", 0.0) for d in items) or 1.0
    summary = {}
    for d in items:
        s = summary.setdefault(d["kernel"], {"launches": 0, "_t": 0.0, "_r": 0.0, "_w": 0.0})
        s["launches"] += 1
        s["_t"] += d.get("time_us", 0.0)
        s["_r"] += d.get("dram_read_bytes", 0.0)
        s["_w"] += d.get("dram_write_bytes", 0.0)
    out = {k: {"launches": s["launches"], "mean_us": s["_t"] / s["launches"], "share_of_profiled_time": s["_t"] / total,
               "dram_read_bytes_mean": s["_r"] / s["launches"], "dram_write_bytes_mean": s["_w"] / s["launches"]}
           for k, s in summary.items()}
    json.dump(out, open(out_summary, "w"), indent=1)
    print(f"{len(items)} launches, {len(out)} kernels -> {out_jsonl}, {out_summary}")


def full(raw_csv, out_csv, header=""):
    rows = csv_rows(raw_csv)
    head, units = rows[0], rows[1]
    launches_ = [r for r in rows[2:] if len(r) == len(head)]
    col = {name: i for i, name in enumerate(head)}
    names = [r[col["Kernel Name"]].split("(")[0] for r in launches_]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f)
        if header:
            w.writerow(["# " + header])
        w.writerow(["metric", "unit"] + names)
        for fixed in ("Kernel Name", "Grid Size", "Block Size"):
            w.writerow([fixed, ""] + [r[col[fixed]] for r in launches_])
        for m in KEEP:
            if m in col:
                w.writerow([m, units[col[m]]] + [r[col[m]] for r in launches_])
    print(f"{len(launches_)} launches x {sum(m in col for m in KEEP)} metrics -> {out_csv}")


if __name__ == "__main__":
    if len(sys.argv) >= 5 and sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "")
    elif len(sys.argv) >= 4 and sys.argv[1] == "full":
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        raise SystemExit(__doc__)
