"""Summarise one `ncu --set full` capture (.ncu-rep, first kernel) into the JSON shape profiles/ uses.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json "note"
Reads the report with `ncu -i ... --page raw --csv` (no GPU needed)."""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "sm__icc_request_hit_rate.pct", "sm__inst_executed.avg.per_cycle_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def main():
    rep, out, note = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {"kernel": vals[hdr.index("Kernel Name")]}
    stalls = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            d[h] = {"value": v, "unit": u}
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            try:
                x = float(v)
            except ValueError:
                continue
            if x >= 0.05:
                stalls[h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = round(x, 6)
    d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    d["note"] = note
    json.dump(d, open(out, "w"), indent=1)
    print(out, d["gpu__time_duration.sum"], d["sm__inst_executed.avg.per_cycle_elapsed"]["value"])


if __name__ == "__main__":
    main()
