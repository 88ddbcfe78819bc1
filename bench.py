#!/usr/bin/env python
"""Headline benchmark: VATEX-large-shape CARE, beam-5 caption decode (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
    python bench.py --impl reference ...                     (CPU arm: the oracle port of the reference)

One "step" = one `translate_batch` over one batch of synthetic videos: encoder + concept head +
29 beam-search decode steps + hypothesis extraction (+ the NCCL all-gather of the decoded ids when
N > 1).  `value` times it with the feature tensors already resident in HBM; `e2e` times the same call
through the public Translator API from pinned HOST feature buffers to host Python lists.  Videos are
independent units, so the batch is sharded over the ranks with a fixed per-GPU batch ("weak").
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "captions_per_sec"
UNIT = "captions/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="care", choices=["care", "reference"])
    ap.add_argument("--config", default="cfg4")
    ap.add_argument("--batch", type=int, default=4096, help="videos per GPU per step")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 (default) / bf16: 16-bit tensor-core operands, fp32 accumulation; fp32: the bit-exact parity mode")
    ap.add_argument("--cpu-batch", type=int, default=16, help="videos per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--self-compact", type=int, default=None, choices=(0, 1, 2, 3),
                    help="bf16 self-attention: 0 = dense tiles, 1 = per-CTA gather of the live cache slots, "
                         "2 = gathered chunk stream (the library default)")
    ap.add_argument("--no-latency", action="store_true", help="skip the small-batch latency section (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-fed e2e section (profiling runs)")
    return ap.parse_args()


def workload_name(cfg, batch, precision):
    return "%s: VATEX-shape Transformer large CARE (d=1024,H=16,F=4096,V=14745,Lm=114), beam 5, 29 steps, " \
           "%d videos/GPU, %s" % (cfg, batch, precision) if cfg == "cfg4" else "%s batch %d %s" % (cfg, batch, precision)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, on the host cores
# ------------------------------------------------------------------------------------------------
def time_cpu_oracle(cfg, cpu_batch, steps, warmup):
    import torch
    from oracle import care_oracle as co
    from oracle.shapes import CONFIGS, make_feats, make_opt
    from oracle.weights import make_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    opt = make_opt(**CONFIGS[cfg])
    sd = make_state_dict(opt, seed=0)
    feats = make_feats(opt, cpu_batch, seed=0)
    for _ in range(warmup):
        co.translate(sd, opt, feats)
    t0 = time.perf_counter()
    for _ in range(steps):
        co.translate(sd, opt, feats)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return cpu_batch / dt, dt * 1e3, cores


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    value, ms, cores = time_cpu_oracle(args.config, args.cpu_batch, steps, warmup)
    sample = "%d videos per step (full 29-step beam-5 decode), %d timed steps, oracle port of the reference's " \
             "CPU path, fp32, %d torch threads on %s" % (args.cpu_batch, steps, cores, cpu_model_name())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.cpu_batch, "fp32 (CPU)"), "beam_size": 5},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# per-kernel CUDA-event timing on the launching stream (roofline evidence)
# ------------------------------------------------------------------------------------------------
class TimedLib:
    """Proxy of the ctypes library: brackets chosen C-ABI calls with CUDA events recorded on torch's
    current stream - the stream the engine passes to the library, i.e. the launching stream."""

    def __init__(self, lib, work):
        self._lib, self._work, self.on, self.records = lib, work, False, []

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name not in self._work:
            return fn
        import torch

        def wrapped(*args):
            if not self.on:
                return fn(*args)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.records.append((name, self._work[name](args), e0, e1))
            return rc
        return wrapped


def kernel_work_table(esz):
    """name -> f(args) = (kernel label, bound, algorithmic units of one launch); DESIGN.md section 4."""
    return {
        "care_gemm": lambda a: ("gemm_bf16_tcgen05_kernel / gemm_bf16_2sm_kernel", "tensor", 2.0 * a[10] * a[11] * a[12]),
        "care_vocab_beam_partials": lambda a: ("vocab_beam_tcgen05_kernel", "tensor", 2.0 * a[5] * a[6] * a[7]),
        # cross: K/V of every video once + q in + ctx out
        "care_cross_attn_step": lambda a: ("attn_mma_kernel<cross>", "hbm",
                                           (a[6] * a[5] * 2.0 * a[9] + 2.0 * a[6] * a[7] * a[9]) * esz),
        # self at step t: the K*t cached keys/values of every video once + q in + ctx out
        # (all K slots of every position; run_care_arm replaces the K/V part by the rows the kernels actually
        # read, counted on the device, when the live-slot stream kernel is in use)
        "care_self_attn_step": lambda a: (SELF_LABEL, "hbm",
                                          (a[4] * a[5] * a[3] * 2.0 * a[7] + 2.0 * a[4] * a[5] * a[7]) * esz),
        "care_add_ln": lambda a: ("add_ln_kernel", "hbm", a[7] * a[8] * (4.0 + 2 * esz)),
        "care_beam_step": lambda a: ("beam_row_kernel", "hbm", 0.0),
    }


SELF_LABEL = "attn_self_stream_kernel / attn_mma_kernel<self>"


def summarise_kernels(records, peaks, units_override=None):
    import collections
    agg = collections.OrderedDict()
    for name, (label, bound, units), e0, e1 in records:
        ms = e0.elapsed_time(e1)
        d = agg.setdefault(label, dict(bound=bound, ms=0.0, units=0.0, n=0))
        d["ms"] += ms
        d["units"] += units
        d["n"] += 1
    for label, total in (units_override or {}).items():
        if label in agg:
            agg[label]["units"] = total
    hbm = peaks.get("hbm_gbs") or 6650.0
    tens = peaks.get("bf16_tflops_sustained") or 1400.0
    src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
    out = []
    for label, d in agg.items():
        if d["ms"] <= 0 or d["units"] <= 0:
            continue
        if d["bound"] == "hbm":
            ach, peak, unit = d["units"] / (d["ms"] * 1e-3) / 1e9, hbm, "GB/s"
            psrc = src + " hbm_gbs (measured copy bandwidth)"
        else:
            ach, peak, unit = d["units"] / (d["ms"] * 1e-3) / 1e12, tens, "TFLOP/s"
            psrc = src + " bf16_tflops_sustained (kernel timed inside a long step)"
        out.append({"kernel": label, "bound": d["bound"], "achieved": ach, "peak": peak, "unit": unit,
                    "frac": ach / peak, "traffic": None, "launches_timed": d["n"],
                    "avg_launch_ms": d["ms"] / d["n"], "total_ms": d["ms"],
                    "algorithmic_units_per_launch": d["units"] / d["n"], "peak_source": psrc})
    out.sort(key=lambda r: -r["total_ms"])
    return out


def read_counter(eng, name):
    import ctypes
    v = ctypes.c_int64(0)
    rc = eng.lib.care_ctx_counter(eng.ctx, name.encode(), ctypes.byref(v))
    return int(v.value) if rc == 0 else 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_care_arm(args):
    import torch
    import torch.distributed as dist
    import care_b200
    from care_b200 import sharding
    from oracle.shapes import CONFIGS, make_feats, make_opt      # synthetic shapes/weights only
    from oracle.weights import make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    opt = make_opt(**CONFIGS[args.config])
    sd = make_state_dict(opt, seed=0)
    model = care_b200.get_framework(dict(opt, care_precision=args.precision, care_self_compact=args.self_compact))
    # (care_self_compact None keeps the library default)
    model.load_state_dict(sd)
    model = model.eval().to(dev)
    del sd
    tr = care_b200.get_translator(opt)
    eng = model.engine()
    B = args.batch
    Tm = opt["max_len"] - 1
    # distinct videos per rank: seeds differ, generated in chunks to bound host memory
    chunks = []
    for c in range(0, B, 512):
        n = min(512, B - c)
        chunks.append(make_feats(opt, n, seed=1000 * rank + c))
    host_feats = [torch.cat([ch[i] for ch in chunks]).pin_memory() for i in range(len(opt["modality"]))]
    del chunks
    dev_feats = [f.to(dev) for f in host_feats]
    h2d_bytes = sum(f.numel() * f.element_size() for f in host_feats)

    def step_resident():
        out = tr.decode_on_device(model, dev_feats)
        if world > 1:  # the one collective of the path: all-gather of the decoded ids
            return sharding.gather_hypotheses(sharding.pack_hypotheses(*out), world * B)
        return out

    # N > 1: every rank ends a step holding (a) the all-gathered ids of ALL videos, read back to a pinned host
    # tensor, and (b) Python lists for its own shard (what a per-rank caption writer consumes)
    gathered_host = torch.empty((world * B, 1, Tm + 3), dtype=torch.int32).pin_memory() if world > 1 else None

    def gather_hook(out):
        full = sharding.gather_hypotheses(sharding.pack_hypotheses(*out), world * B)
        gathered_host.copy_(full, non_blocking=True)
        return out

    def step_e2e():
        if world > 1:
            with torch.no_grad():
                if B > tr.pipeline_chunk:
                    out = tr.decode_pipelined(model, host_feats, tr.pipeline_chunk)
                else:
                    out = tr.decode_on_device(model, host_feats)
            out = gather_hook(out)
            return care_b200.engine.hyps_from_device(*out, tr.beam_alpha, tr.topk)
        return tr.translate_batch([model], {"feats": host_feats})

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    esz = 4 if args.precision == "fp32" else 2
    timed = TimedLib(eng.lib, kernel_work_table(esz))
    eng.lib = timed

    for _ in range(args.warmup):
        step_resident()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    timed.on = True
    launches0 = eng.launch_count()
    rows0 = read_counter(eng, "self_attn_rows")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    sync_all()
    timed.on = False
    launches = eng.launch_count() - launches0
    self_rows = read_counter(eng, "self_attn_rows") - rows0   # K/V cache rows the bf16 self-attention kernels read
    elapsed_ms = ev0.elapsed_time(ev1)
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms / 1e3)

    # e2e: host pinned features -> H2D -> decode -> (all-gather) -> D2H -> Python lists, public API.
    # (a) one synchronous Translator.translate_batch call per step;
    # (b) Translator.translate_stream over the same steps: every step's H2D copy and D2H read are inside the
    #     timed region, but step i+1's copy overlaps step i's decode (what a loader loop gets).
    e2e_value = e2e_call_value = None
    e2e_steps = e2e_stream_steps = 0
    hyps = None
    if not args.no_e2e:
        for _ in range(1):
            step_e2e()
        sync_all()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            hyps, scores = step_e2e()
        torch.cuda.synchronize(dev)
        e2e_call_ms = (time.perf_counter() - t0) * 1e3
        t = torch.tensor([e2e_call_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_call_value = world * B * e2e_steps / (float(t.item()) / 1e3)

        def stream_steps(n):
            got = 0
            hook = gather_hook if world > 1 else None
            for h, s_ in tr.translate_stream([model], ({"feats": host_feats} for _ in range(n)), device_hook=hook):
                got += len(h)
            return got

        stream_steps(2)
        sync_all()
        # a loader loop runs many batches; 16 keeps the one-off pipeline fill (first H2D, last read-back) in proportion
        e2e_stream_steps = max(args.steps, 16)
        t0 = time.perf_counter()
        got = stream_steps(e2e_stream_steps)
        torch.cuda.synchronize(dev)
        e2e_ms = (time.perf_counter() - t0) * 1e3
        assert got == B * e2e_stream_steps
        t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = world * B * e2e_stream_steps / (float(t.item()) / 1e3)
    clocks = sampler.stop() if rank == 0 else None
    d2h_bytes = B * world * (Tm + 3) * 4
    assert hyps is None or (len(hyps) == B and all(1 <= len(h[0]) <= Tm for h in hyps[:64]))
    if hyps is not None and world > 1:   # the gathered record of this rank's first video matches its own list
        lo = rank * B
        assert gathered_host[lo, 0, :len(hyps[0][0])].tolist() == hyps[0][0]

    # per-step decode latency at small batches (launch-bound regime: the decode is replayed as one CUDA graph)
    latency = {}
    if world == 1 and not args.no_latency:
        timed.on = False
        for lb in (1, 64):
            small = [f[:lb].contiguous() for f in dev_feats]
            for _ in range(3):
                tr.decode_on_device(model, small)
            torch.cuda.synchronize(dev)
            reps = 10
            l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0.record()
            for _ in range(reps):
                tr.decode_on_device(model, small)
            l1.record()
            torch.cuda.synchronize(dev)
            ms = l0.elapsed_time(l1) / reps
            latency["batch_%d" % lb] = {"ms_per_caption_batch": ms, "us_per_beam_step": ms / Tm * 1e3,
                                        "captions_per_sec": lb / ms * 1e3}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    override = {}
    if self_rows > 0:   # algorithmic bytes of the self-attention = the rows actually read + q in + ctx out
        n_self = sum(1 for r in timed.records if r[0] == "care_self_attn_step")
        override[SELF_LABEL] = (self_rows * 2.0 * eng.d + n_self * 2.0 * B * opt["beam_size"] * eng.d) * esz
    kernels = summarise_kernels(timed.records, peaks, override)
    # DRAM traffic per launch from the committed `ncu --set full` capture of one decode step (profiles/)
    import glob
    tfiles = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic_*.json")))
    if tfiles and args.config == "cfg4" and B == 4096 and args.precision != "fp32":
        tr_json = json.load(open(tfiles[-1]))
        for r in kernels:
            tkey = r["kernel"].split(" / ")[0]
            if tkey in tr_json:
                r["traffic"] = tr_json[tkey]["traffic_bytes_per_launch"]
                r["traffic_source"] = "%s (dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches " \
                                      "of one decode step, t~15)" % os.path.relpath(tfiles[-1], ROOT)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cv, cms, cores = time_cpu_oracle(args.config, args.cpu_batch, 3, 1)
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "%d videos per pass, 3 timed passes of the full 29-step beam-5 decode (%.1f s each) after one "
                         "warm-up, oracle port of the reference CPU path, fp32, %d torch threads on %s" % (
                             args.cpu_batch, cms / 1e3, cores, cpu_model_name())}
    step_ms = elapsed_ms / args.steps
    roofline = dict(kernels[0]) if kernels else None
    if roofline is not None:
        roofline["share_of_step"] = roofline["total_ms"] / elapsed_ms
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": workload_name(args.config, B, args.precision), "beam_size": opt["beam_size"],
                   "per_gpu_batch": B, "global_batch": B * world, "parallelism": "video-sharded x%d" % world,
                   "decode_ms_per_beam_step": step_ms / Tm,
                   "small_batch_latency": latency,
                   "l2_note": "inputs larger than L2: per-step working set (cross K/V 1.9 GB, KV cache up to 3.6 GB, "
                              "features 1.4 GB) >> 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_stream_steps or e2e_steps,
                "single_call_value": e2e_call_value,
                "note": "value: Translator.translate_stream over the steps' pinned HOST batches -> Python lists "
                        "(every step's H2D copy and D2H read inside the timed region; step i+1's copy overlaps step "
                        "i's decode; N>1: + the per-step NCCL all-gather of ids, read back in full to pinned host memory on every rank, "
                        "Python lists built for the rank's own shard).  single_call_value: one synchronous "
                        "Translator.translate_batch per step (H2D chunked in 2048-video halves)"},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
        "roofline_other_kernels": [{k: r[k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac", "traffic",
                                                       "launches_timed", "avg_launch_ms", "total_ms",
                                                       "algorithmic_units_per_launch")}
                                   for r in kernels[1:]],
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The ONE JSON line goes to the process's original stdout."""
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


_JSON_OUT = sys.stdout


def main():
    global _JSON_OUT
    args = parse()
    # libraries that write to file descriptor 1 (NCCL prints its version banner there) must not end up in
    # front of the JSON line: keep a private handle on the real stdout and point fd 1 at stderr
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_care_arm(args)


if __name__ == "__main__":
    main()
