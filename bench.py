#!/usr/bin/env python
"""Headline benchmark: VATEX-large-shape CARE, beam-5 caption decode (BASELINE.json configs[3]).

    python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
    python bench.py --impl reference ...                     (CPU arm: the oracle port of the reference)

One "step" = one `translate_batch` over one batch of synthetic videos: encoder + concept head +
29 beam-search decode steps + hypothesis extraction (+ the NCCL all-gather of the decoded ids when
N > 1).  `value` times it with the feature tensors already resident in HBM; `e2e` times the same call
through the public Translator API from pinned HOST feature buffers to host Python lists.  Videos are
independent units: the GLOBAL batch of 4096 videos (BASELINE.json configs[3]: "batch 4096 sharded across
1/2/4/8 B200") is split over the ranks - "scaling": "strong"; the weak-scaling number (4096 videos on every
GPU) rides along as `config.other_scaling` when N > 1.  `--config cfg5` runs BASELINE.json configs[4]
(mask-predict, global batch 1024).  Prints ONE JSON line on rank 0.
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
    ap.add_argument("--batch", type=int, default=0,
                    help="videos per step: the GLOBAL batch (strong scaling) or the per-GPU batch (weak); "
                         "default 4096 (cfg4) / 1024 (cfg5), BASELINE.json configs[3] / configs[4]")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): one global batch sharded over the GPUs, as BASELINE.json configs[3] "
                         "words it; weak: --batch videos on every GPU")
    ap.add_argument("--no-other-scaling", action="store_true", help="N > 1: skip the secondary (weak) record")
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "fp32"],
                    help="fp16 (default) / bf16: 16-bit tensor-core operands, fp32 accumulation; fp32: the bit-exact parity mode")
    ap.add_argument("--cpu-batch", type=int, default=16, help="videos per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--self-compact", type=int, default=None, choices=(0, 1, 2, 3),
                    help="bf16 self-attention: 0 = dense tiles, 1 = per-CTA gather of the live cache slots, "
                         "2 = gathered chunk stream (the library default)")
    ap.add_argument("--graph-lanes", type=int, default=None,
                    help="concurrent lanes of a graph-replayed decode (A/B runs; default: the engine's automatic choice)")
    ap.add_argument("--graph-max-rows", type=int, default=None,
                    help="largest number of beam rows decoded as one CUDA graph (A/B runs; engine default 6144)")
    ap.add_argument("--no-latency", action="store_true", help="skip the small-batch latency section (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-fed e2e section (profiling runs)")
    return ap.parse_args()


def workload_name(cfg, n_global, world, precision, scaling="strong"):
    how = "global batch %d sharded over %d GPU(s)" % (n_global, world) if scaling == "strong" else \
        "%d videos on each of %d GPU(s)" % (n_global // max(world, 1), world)
    if cfg == "cfg4":
        return "cfg4: VATEX-shape Transformer large CARE (d=1024,H=16,F=4096,V=14745,Lm=114), beam 5, 29 steps, " \
               "%s, %s" % (how, precision)
    if cfg == "cfg5":
        return "cfg5: MSRVTT-shape NACF CARE (ARB encoder, d=512,V=10547,Lm=114), mask-predict: 6 length " \
               "candidates, coarse templates + 5 iterations, %s, %s" % (how, precision)
    return "%s, %s, %s" % (cfg, how, precision)


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference, on the host cores
# ------------------------------------------------------------------------------------------------
def time_cpu_oracle(cfg, cpu_batch, steps, warmup):
    """The reference's CPU path on the host cores: the UNMODIFIED reference (oracle/_ref, vendored by
    oracle/build_ref.py; kind "reference") when it is present, else the oracle restatement (kind "port").
    Returns (captions/s, ms per pass, threads, kind)."""
    import torch
    from oracle import care_oracle as co
    from oracle import ref_harness as rh
    from synth.shapes import CONFIGS, make_feats, make_opt
    from synth.weights import make_state_dict
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    opt = make_opt(**CONFIGS[cfg])
    sd = make_state_dict(opt, seed=0, perturb=opt["decoding_type"] == "NARFormer")
    feats = make_feats(opt, cpu_batch, seed=0)
    kind = "port"
    run = lambda: co.translate(sd, opt, feats)   # noqa: E731
    if rh.reference_available():
        try:
            model = rh.build_reference_model(opt)
            model.load_state_dict(sd, strict=True)
            run = lambda: rh.run_reference_translate(model, opt, [f.clone() for f in feats])   # noqa: E731
            run()
            kind = "reference"
        except Exception as exc:   # e.g. a partial copy: fall back to the restatement and say so
            sys.stderr.write("[bench] reference not runnable (%r): timing the oracle port\n" % (exc,))
            run = lambda: co.translate(sd, opt, feats)   # noqa: E731
    for _ in range(warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(steps):
        run()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return cpu_batch / dt, dt * 1e3, cores, kind


CPU_KIND_TEXT = {"reference": "the unmodified reference's own Translator (oracle/_ref) on the CPU",
                 "port": "oracle port of the reference's CPU path"}


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
    value, ms, cores, kind = time_cpu_oracle(args.config, args.cpu_batch, steps, warmup)
    sample = "%d videos per step (full decode), %d timed steps, %s, fp32, %d torch threads on %s" % (
        args.cpu_batch, steps, CPU_KIND_TEXT[kind], cores, cpu_model_name())
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, args.cpu_batch, 1, "fp32 (CPU, %d-video sample per step)"
                                             % args.cpu_batch), "beam_size": 5},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
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


def kernel_work_table(esz, lib, ctx):
    """name -> f(args) = (kernel label, bound, algorithmic units of one launch); DESIGN.md section 4.  The label of
    a family with several kernel variants is the variant the call actually launched (care_ctx_last_kernel)."""
    def last(family, default):
        def f():
            n = lib.care_ctx_last_kernel(ctx, family.encode())
            return n.decode() if n else default
        return f
    gemm, vocab, selfk = last("gemm", "gemm"), last("vocab", "vocab_beam"), last("self_attn", "self_attn")
    return {
        "care_gemm": lambda a: (gemm(), "tensor", 2.0 * a[10] * a[11] * a[12]),
        "care_vocab_beam_partials": lambda a: (vocab(), "tensor", 2.0 * a[5] * a[6] * a[7]),
        # cross: K/V of every video once + q in + ctx out
        "care_cross_attn_step": lambda a: ("attn_mma_kernel<cross>", "hbm",
                                           (a[6] * a[5] * 2.0 * a[9] + 2.0 * a[6] * a[7] * a[9]) * esz),
        # self at step t: the K*t cached keys/values of every video once + q in + ctx out
        # (all K slots of every position; run_care_arm replaces the K/V part by the rows the kernels actually
        # read, counted on the device, when the live-slot stream kernel is in use)
        "care_self_attn_step": lambda a: (SELF_LABEL, "hbm",
                                          (a[4] * a[5] * a[3] * 2.0 * a[7] + 2.0 * a[4] * a[5] * a[7]) * esz),
        # full-sequence attention (mask-predict passes): per group K and V once + q in + ctx out
        "care_group_attn": lambda a: ("group_attn_mma_kernel", "hbm", a[8] * (a[10] * 2.0 + a[9] * 2.0) * a[12] * esz),
        # args: ctx, A, lda, W, ldw, bias, residual, residual_dtype, gamma, beta, eps, out16, out32, M, N, K, stream
        "care_gemm_add_ln": lambda a: (gemm(), "tensor", 2.0 * a[13] * a[14] * a[15]),
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
            psrc = src + " bf16_tflops_sustained (16-bit dense tensor peak; kernel timed inside a long step)"
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


def percentile(xs, q):
    xs = sorted(xs)
    if not xs:
        return None
    k = (len(xs) - 1) * q
    lo, hi = int(k), min(int(k) + 1, len(xs) - 1)
    return xs[lo] + (xs[hi] - xs[lo]) * (k - lo)


def make_host_feats(opt, n, seed_base):
    """Pinned host features of n distinct videos, generated in chunks to bound host memory."""
    import torch
    from synth.shapes import make_feats
    chunks = []
    for c in range(0, max(n, 1), 512):
        chunks.append(make_feats(opt, min(512, n - c), seed=seed_base + c))
    return [torch.cat([ch[i] for ch in chunks]).pin_memory() for i in range(len(opt["modality"]))]


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_care_arm(args):
    import torch
    import torch.distributed as dist
    import care_b200
    from care_b200 import sharding
    from synth.shapes import CONFIGS, make_opt      # synthetic shapes / weights: pure data builders
    from synth.weights import make_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    opt = make_opt(**CONFIGS[args.config])
    nar = opt["decoding_type"] == "NARFormer"
    sd = make_state_dict(opt, seed=0, perturb=nar)
    extra = {} if args.graph_lanes is None else {"care_graph_lanes": args.graph_lanes}
    if args.graph_max_rows is not None:
        extra["care_cuda_graph_max_rows"] = args.graph_max_rows
    model = care_b200.get_framework(dict(opt, care_precision=args.precision, care_self_compact=args.self_compact, **extra))
    # (care_self_compact None keeps the library default)
    model.load_state_dict(sd)
    model = model.eval().to(dev)
    del sd
    tr = care_b200.get_translator(opt)
    eng = model.engine()
    Tm = opt["max_len"] - 1
    K = 1 if nar else opt["beam_size"]
    G = args.batch if args.batch else (1024 if nar else 4096)          # videos per step
    if args.scaling == "strong":      # BASELINE.json configs[3]/[4]: ONE global batch sharded over the GPUs
        lo, hi = sharding.shard_range(G, rank, world)
        B, n_global = hi - lo, G
    else:                             # fixed per-GPU batch
        B, n_global = G, G * world
    host_feats = make_host_feats(opt, B, 100000 * rank)
    dev_feats = [f.to(dev) for f in host_feats]
    h2d_bytes = sum(f.numel() * f.element_size() for f in host_feats)

    def pack(out):
        return sharding.pack_nar(*out, opt["max_len"]) if nar else sharding.pack_hypotheses(*out)

    def step_resident(feats=None, total=None):
        out = tr.decode_on_device(model, dev_feats if feats is None else feats)
        if world > 1:  # the one collective of the path: all-gather of the decoded ids
            return sharding.gather_hypotheses(pack(out), n_global if total is None else total)
        return out

    # N > 1: every rank ends a step holding (a) the all-gathered ids of ALL videos, read back to a pinned host
    # tensor, and (b) Python lists for its own shard (what a per-rank caption writer consumes)
    rec_w = 2 * opt["max_len"] + 1 if nar else Tm + 3
    gathered_host = torch.empty((n_global, 1, rec_w), dtype=torch.int32).pin_memory() if world > 1 else None

    def gather_hook(out):
        full = sharding.gather_hypotheses(pack(out), n_global)
        gathered_host.copy_(full, non_blocking=True)
        return out

    def step_e2e():
        if nar:
            with torch.no_grad():
                out = tr.decode_on_device(model, host_feats)
            if world > 1:
                out = gather_hook(out)
            return out[0].cpu().tolist(), out[1].cpu().tolist()
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

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    esz = 4 if args.precision == "fp32" else 2
    raw_lib = eng.lib
    timed = TimedLib(raw_lib, kernel_work_table(esz, raw_lib, eng.ctx))
    eng.lib = timed

    for _ in range(args.warmup):
        step_resident()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident()
    ev1.record()
    sync_all()
    launches = eng.launch_count() - launches0
    elapsed_ms = max_over_ranks(ev0.elapsed_time(ev1))
    value = n_global * args.steps / (elapsed_ms / 1e3)
    # Roofline evidence: the timed steps replay one CUDA graph per batch, which CUDA events cannot bracket, so the
    # SAME steps are repeated eagerly right away with an event pair around every C-ABI launch of the dominant kernels
    # (on the launching stream); kernel shares are taken against this instrumented pass's own duration.
    use_graphs = eng.use_graphs
    eng.use_graphs = False
    step_resident()
    sync_all()
    timed.on = True
    rows0 = read_counter(eng, "self_attn_rows")
    iv0, iv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iv0.record()
    for _ in range(args.steps):
        step_resident()
    iv1.record()
    sync_all()
    timed.on = False
    eng.use_graphs = use_graphs
    self_rows = read_counter(eng, "self_attn_rows") - rows0   # K/V cache rows the self-attention kernels read
    instr_ms = max_over_ranks(iv0.elapsed_time(iv1))

    # e2e: host pinned features -> H2D -> decode -> (all-gather) -> D2H -> Python lists, public API.
    # (a) one synchronous Translator.translate_batch call per step;
    # (b) Translator.translate_stream over the same steps: every step's H2D copy and D2H read are inside the
    #     timed region, but step i+1's copy overlaps step i's decode (what a loader loop gets).
    e2e_value = e2e_call_value = None
    e2e_steps = e2e_stream_steps = 0
    hyps = None
    stages = {}
    if not args.no_e2e:
        for _ in range(2):   # the second decode of a shape captures its CUDA graph (engine: graph_max_rows_repeat)
            step_e2e()
        sync_all()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            hyps, scores = step_e2e()
        torch.cuda.synchronize(dev)
        e2e_call_value = n_global * e2e_steps / (max_over_ranks((time.perf_counter() - t0) * 1e3) / 1e3)
        if nar:     # mask-predict has no stream API: the synchronous call is the e2e number
            e2e_value, e2e_stream_steps = e2e_call_value, e2e_steps
        else:
            def stream_steps(n):
                got = 0
                hook = gather_hook if world > 1 else None
                marks = [time.perf_counter()]
                for h, s_ in tr.translate_stream([model], ({"feats": host_feats} for _ in range(n)), device_hook=hook):
                    got += len(h)
                    marks.append(time.perf_counter())
                if os.environ.get("CARE_B200_DEBUG") and rank == 0:
                    iv = ["%.2f" % ((b - a) * 1e3) for a, b in zip(marks, marks[1:])]
                    sys.stderr.write("[bench] stream of %d: ms between results: %s\n" % (n, " ".join(iv)))
                return got

            # a loader loop runs many batches; 32 keeps the one-off pipeline fill (the first H2D copy and the last
            # read-back + list building, ~32 ms together at 4096 videos) at ~1 ms per step
            e2e_stream_steps = max(args.steps, 32)
            # warm-up: one stream of the same length (scripts/stream_probe.py: the first LONG stream of a process pays
            # ~100 ms of pinned / staging allocations and ~40 ms more at its end, a 3-batch one does not pay all of it;
            # later streams run at the resident rate + the pipeline fill)
            stream_steps(e2e_stream_steps)
            sync_all()
            t0 = time.perf_counter()
            got = stream_steps(e2e_stream_steps)
            torch.cuda.synchronize(dev)
            e2e_ms = (time.perf_counter() - t0) * 1e3
            assert got == B * e2e_stream_steps
            e2e_value = n_global * e2e_stream_steps / (max_over_ranks(e2e_ms) / 1e3)
        if world == 1 and not nar:
            # the stages of one host-fed batch, each timed alone (in the stream they overlap)
            stage_dev = [torch.empty_like(f) for f in dev_feats]
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for d_, s_ in zip(stage_dev, host_feats):
                d_.copy_(s_, non_blocking=True)
            torch.cuda.synchronize(dev)
            stages["h2d_ms"] = (time.perf_counter() - t0) * 1e3
            out = tr.decode_on_device(model, dev_feats, early_exit_every=0)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            host_out = [t.cpu() for t in out]
            stages["d2h_ms"] = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter()
            care_b200.engine.hyps_from_device(*host_out, tr.beam_alpha, tr.topk)
            stages["lists_ms"] = (time.perf_counter() - t0) * 1e3
            stages["decode_ms"] = elapsed_ms / args.steps
            del stage_dev
    clocks = sampler.stop() if rank == 0 else None
    d2h_bytes = n_global * rec_w * 4
    if hyps is not None and not nar:
        assert len(hyps) == B and all(1 <= len(h[0]) <= Tm for h in hyps[:64])
        if world > 1 and B:   # the gathered record of this rank's first video matches its own list
            lo = sharding.shard_range(n_global, rank, world)[0] if args.scaling == "strong" else rank * B
            assert gathered_host[lo, 0, :len(hyps[0][0])].tolist() == hyps[0][0]

    # per-step decode latency: every beam step of the eager decode bracketed by CUDA events (this rank's shard),
    # and whole-decode latency at small batches (launch-bound regime: replayed as one CUDA graph)
    latency = {}
    step_lat = None
    if not args.no_latency and not nar:
        timed.on = False
        step_ms = []
        enc = model.encoding_phase(dev_feats)
        orig = eng.decode_step

        def timed_step(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = orig(*a, **k)
            e1.record()
            step_ms.append((e0, e1))
            return r

        eng.decode_step = timed_step
        use_graphs = eng.use_graphs
        eng.use_graphs = False
        try:
            for _ in range(3):
                eng.ar_decode(enc, B, beam_size=tr.beam_size, topk=tr.topk, beam_alpha=tr.beam_alpha, early_exit_every=0)
        finally:
            eng.decode_step, eng.use_graphs = orig, use_graphs
        torch.cuda.synchronize(dev)
        per = [a.elapsed_time(b) for a, b in step_ms[Tm:]]     # first decode = warm-up
        step_lat = {"videos": B, "beam_rows": B * K, "samples": len(per), "mean_ms": sum(per) / len(per),
                    "p50_ms": percentile(per, 0.5), "p99_ms": percentile(per, 0.99), "max_ms": max(per),
                    "note": "CUDA events around each of the %d beam steps (10-11 layer launches + fused vocabulary + "
                            "beam update) of the eager decode, 2 decodes after one warm-up" % Tm}
        if world == 1:
            for lb in (1, 64, 512):
                if lb > B:
                    continue
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

    # the other scaling mode as a secondary record (N > 1): weak = 4096 videos on every GPU
    other = None
    if world > 1 and not args.no_other_scaling:
        if args.scaling == "strong":
            ob, ototal = G, G * world
        else:
            olo, ohi = sharding.shard_range(G, rank, world)
            ob, ototal = ohi - olo, G
        ofe = [f.to(dev) for f in make_host_feats(opt, ob, 100000 * rank + 7)]
        for _ in range(2):
            step_resident(ofe, ototal)
        sync_all()
        osteps = max(2, args.steps // 2)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        o0.record()
        for _ in range(osteps):
            step_resident(ofe, ototal)
        o1.record()
        sync_all()
        oms = max_over_ranks(o0.elapsed_time(o1))
        other = {"scaling": "weak" if args.scaling == "strong" else "strong", "global_batch": ototal,
                 "per_gpu_batch": ob, "steps": osteps, "ms_per_step": oms / osteps,
                 "value": ototal * osteps / (oms / 1e3), "unit": UNIT}
        del ofe

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    override = {}
    if self_rows > 0:   # algorithmic bytes of the self-attention = the rows actually read + q in + ctx out
        n_self = sum(1 for r in timed.records if r[0] == "care_self_attn_step")
        override[SELF_LABEL] = (self_rows * 2.0 * eng.d + n_self * 2.0 * B * K * eng.d) * esz
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
        cv, cms, cores, ckind = time_cpu_oracle(args.config, args.cpu_batch, 3, 1)
        cpu = {"value": cv, "unit": UNIT, "cores": cores, "kind": ckind,
               "sample": "%d videos per pass, 3 timed passes of the full decode (%.1f s each) after one "
                         "warm-up, %s, fp32, %d torch threads on %s" % (
                             args.cpu_batch, cms / 1e3, CPU_KIND_TEXT[ckind], cores, cpu_model_name())}
    step_ms_avg = elapsed_ms / args.steps
    roofline = dict(kernels[0]) if kernels else None
    if roofline is not None:
        roofline["share_of_step"] = roofline["total_ms"] / instr_ms
        roofline["timing"] = ("CUDA events around every launch in an instrumented eager repeat of the timed steps "
                              "(%.3f ms per step; the timed steps replay a CUDA graph: %.3f ms per step)"
                              % (instr_ms / args.steps, elapsed_ms / args.steps))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms_avg, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": workload_name(args.config, n_global, world, args.precision, args.scaling),
                   "beam_size": K, "per_gpu_batch": B, "global_batch": n_global,
                   "parallelism": "video-sharded x%d" % world,
                   "decode_ms_per_beam_step": None if nar else step_ms_avg / Tm,
                   "per_step_latency": step_lat,
                   "small_batch_latency": latency,
                   "other_scaling": other,
                   "l2_note": "inputs larger than L2: per-step working set (cross K/V 1.9 GB, KV cache up to 3.6 GB, "
                              "features 1.4 GB at 4096 videos) >> 126 MB L2"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world,
                "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_stream_steps or e2e_steps,
                "single_call_value": e2e_call_value, "stages_ms": stages,
                "note": "value: Translator.translate_stream over the steps' pinned HOST batches -> Python lists "
                        "(every step's H2D copy and D2H read inside the timed region; step i+1's copy overlaps step "
                        "i's decode; N>1: + the per-step NCCL all-gather of ids, read back in full to pinned host memory on every rank, "
                        "Python lists built for the rank's own shard).  single_call_value: one synchronous "
                        "Translator.translate_batch per step (H2D chunked in 2048-video halves).  stages_ms: the "
                        "stages of one batch timed alone (N=1); in the stream they overlap"},
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
