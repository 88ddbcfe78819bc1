"""Prints the key counters of an .ncu-rep (read on the CPU box): duration, DRAM bytes, throughputs,
occupancy, issue rate, top stall reasons per source line.  Usage: python scripts/ncu_summary.py rep [n_lines]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit", "lts__t_sector_hit_rate.pct", "launch__grid_size",
        "launch__block_size", "sm__maximum_warps_per_active_cycle_pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:90])
        for h, u, v in zip(hdr, units, r):
            if any(h.startswith(k.strip()) if k.endswith(" ") else k in h for k in KEYS):
                print("   %-75s %s %s" % (h, v, u))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[start + 1:] if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    agg = {h: sum(int(r[ix[h]] or 0) for r in data) for h in stalls}
    print("total samples", tot, "instructions", sum(int(r[ix["Instructions Executed"]] or 0) for r in data))
    print("stall totals:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
        st = sorted(((int(r[ix[h]] or 0), h) for h in stalls), reverse=True)[:2]
        print("  %6s %9s  %-70s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Source"]][:70], st))


if __name__ == "__main__":
    main()
