"""One decode step out of an `ncu --set full` capture of bench.py: per-launch table (duration, DRAM bytes,
throughputs, occupancy, L2 hit rate, registers) and the per-kernel DRAM traffic JSON that bench.py attaches to
its roofline lines.  The capture may start anywhere: the first complete step (embed_ln_kernel up to the launch
before the next embed_ln_kernel) is used.
Usage: python scripts/step_profile.py capture.ncu-rep profiles/r01_ncu_full_vN_step.txt profiles/r01_traffic_vN.json "header" """
import collections
import csv
import io
import json
import re
import subprocess
import sys

COLS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "launch__registers_per_thread"]
# labels = the kernel names bench.py reports (care_ctx_last_kernel), so that its roofline records find their traffic
LABELS = [("gemm_add_ln_pair_kernel<0>", "gemm_add_ln_pair_kernel<h16 residual>"),
          ("gemm_add_ln_pair_kernel<1>", "gemm_add_ln_pair_kernel<f32 residual>"),
          (re.compile(r"gemm_add_ln_kernel<0\b"), "gemm_add_ln_kernel<h16 residual>"),
          (re.compile(r"gemm_add_ln_kernel<1\b"), "gemm_add_ln_kernel<f32 residual>"), ("gemm_add_ln_kernel<0>", "gemm_add_ln_kernel<h16 residual>"), ("gemm_add_ln_kernel<1>", "gemm_add_ln_kernel<f32 residual>"),
          ("gemm_bf16_2sm_kernel<float>", "gemm_bf16_2sm_kernel<float>"), ("gemm_bf16_2sm_kernel", "gemm_bf16_2sm_kernel<h16>"),
          (re.compile(r"gemm_bf16_tcgen05_kernel<\d+, float>"), "gemm_bf16_tcgen05_kernel<float>"),
          ("gemm_bf16_tcgen05_kernel", "gemm_bf16_tcgen05_kernel<h16>"),
          ("vocab_beam_2sm_kernel", "vocab_beam_2sm_kernel"), ("vocab_beam_tcgen05_kernel", "vocab_beam_tcgen05_kernel"),
          ("attn_self_stream_kernel", "attn_self_stream_kernel"),
          (re.compile(r"attn_mma_kernel<\d+, 0,"), "attn_mma_kernel<cross>"),
          (re.compile(r"attn_mma_kernel<\d+, 1,"), "attn_mma_kernel<self>"), ("add_ln_kernel", "add_ln_kernel"),
          ("compact_info_kernel", "compact_info_kernel"), ("embed_ln_kernel", "embed_ln_kernel"),
          ("beam_update_kernel", "beam_update_kernel")]


def to_bytes(value, unit):
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value.replace(",", "")) * scale.get(unit, 1.0)


def main():
    rep, out_txt, out_json = sys.argv[1:4]
    header = sys.argv[4] if len(sys.argv) > 4 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    starts = [i for i, r in enumerate(body) if "embed_ln_kernel" in r[name_i]]
    if len(starts) >= 2:
        body = body[starts[0]:starts[1]]
    idx = [hdr.index(c) for c in COLS]
    lines = ["# " + header, "Kernel Name | " + " | ".join(COLS), " | " + " | ".join(units[i] for i in idx)]
    traffic = collections.OrderedDict()
    for r in body:
        lines.append(r[name_i][:64] + " | " + " | ".join(r[i] for i in idx))
        label = next((lab for key, lab in LABELS
                      if (key.search(r[name_i]) if hasattr(key, "search") else key in r[name_i])), None)
        if label is None:
            continue
        nbytes = to_bytes(r[idx[1]], units[idx[1]]) + to_bytes(r[idx[2]], units[idx[2]])
        t = traffic.setdefault(label, {"traffic_bytes_per_launch": 0.0, "launches": 0})
        t["traffic_bytes_per_launch"] += nbytes
        t["launches"] += 1
    for t in traffic.values():
        t["traffic_bytes_per_launch"] /= t["launches"]
    open(out_txt, "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(out_json, "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
