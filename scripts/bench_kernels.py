"""Per-kernel timing at the cfg4 shapes (CUDA events, L2 flushed by working sets >> 126 MB).
Usage: python scripts/bench_kernels.py [attn] [selfc] [beam] [gemm]   (runs on the GPU box)"""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from care_b200 import _lib  # noqa: E402

F32, BF16 = 0, 1
PEAK = 6464.3
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    what = set(sys.argv[1:]) or {"attn", "beam", "gemm"}
    lib = _lib.load()
    h = ctypes.c_void_p()
    _lib.check(lib.care_ctx_create(ctypes.byref(h), 0), "ctx")
    st = torch.cuda.current_stream().cuda_stream
    B, K, H, d, Lm, Tm, V = 4096, 5, 16, 1024, 114, 29, 14745
    R = B * K
    if "attn" in what:
        q = torch.randn(R, d, device="cuda").bfloat16()
        kv = torch.randn(B, Lm, 2 * d, device="cuda").bfloat16()
        bias = torch.randn(H, Lm, device="cuda")
        done = torch.zeros(B, device="cuda", dtype=torch.int32)
        out = torch.zeros(R, d, device="cuda", dtype=torch.bfloat16)
        for impl in (0, 1):
            lib.care_ctx_set_option(h, b"attn_impl", impl)
            ms = timed(lambda: _lib.check(lib.care_cross_attn_step(
                h, BF16, q.data_ptr(), d, kv.data_ptr(), Lm, B, K, H, d, bias.data_ptr(), done.data_ptr(),
                out.data_ptr(), st), "x"))
            nbytes = B * Lm * 2 * d * 2 + 2 * R * d * 2
            print("cross attn impl=%d: %.3f ms  %.0f GB/s (%.1f%% of measured %.0f)" % (
                impl, ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / PEAK, PEAK))
        del kv
        cache = torch.randn(Tm, R, 3 * d, device="cuda").bfloat16()
        anc = torch.randint(0, K, (B, K, Tm), device="cuda", dtype=torch.uint8)
        tok = torch.randint(4, 100, (B, Tm + 1, K), device="cuda", dtype=torch.int32)
        for n_pos in (1, 4, 6, 8, 10, 12, 15, 22, 29):
            for impl in (0, 1):
                lib.care_ctx_set_option(h, b"attn_impl", impl)
                ms = timed(lambda: _lib.check(lib.care_self_attn_step(
                    h, BF16, cache.data_ptr(), n_pos, B, K, H, d, anc.data_ptr(), Tm, tok.data_ptr(),
                    done.data_ptr(), out.data_ptr(), st), "s"))
                nbytes = R * n_pos * 2 * d * 2 + 2 * R * d * 2
                print("self attn n_pos=%2d impl=%d: %.3f ms  %.0f GB/s (%.1f%%)" % (
                    n_pos, impl, ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / PEAK))
        del cache
    if "selfc" in what:
        # self-attention over live slots only, with an ancestry table from simulated beam parents
        # (uniform parent choice: ~2.6 of 5 slots per position stay referenced, like the benchmark weights)
        cache = torch.randn(Tm, R, 3 * d, device="cuda").bfloat16()
        tok = torch.randint(4, 100, (B, Tm + 1, K), device="cuda", dtype=torch.int32)
        done = torch.zeros(B, device="cuda", dtype=torch.int32)
        out = torch.zeros(R, d, device="cuda", dtype=torch.bfloat16)
        g = torch.Generator().manual_seed(0)
        for n_pos in (4, 8, 15, 22, 29):
            anc = torch.zeros(B, K, Tm, dtype=torch.int64)
            for t in range(1, n_pos):      # beam b at step t continues beam parent[b] of step t - 1
                parent = torch.randint(0, K, (B, K), generator=g)
                anc[:, :, :t - 1] = torch.gather(anc[:, :, :t - 1], 1, parent[:, :, None].expand(B, K, t - 1))
                anc[:, :, t - 1] = parent
            live = sum(len(set(anc[v, :, pp].tolist())) for v in range(64) for pp in range(n_pos - 1)) / 64.0 + K
            d_anc = anc.to(torch.uint8).cuda()
            for compact in (0, 1, 2):
                lib.care_ctx_set_option(h, b"self_compact", compact)
                ms = timed(lambda: _lib.check(lib.care_self_attn_step(
                    h, BF16, cache.data_ptr(), n_pos, B, K, H, d, d_anc.data_ptr(), Tm, tok.data_ptr(),
                    done.data_ptr(), out.data_ptr(), st), "s"))
                dense = R * n_pos * 2 * d * 2 + 2 * R * d * 2
                print("self attn n_pos=%2d live rows %.1f of %d, compact=%d: %.3f ms  (dense bytes at %.0f GB/s)" % (
                    n_pos, live, n_pos * K, compact, ms, dense / ms / 1e6))
            lib.care_ctx_set_option(h, b"self_compact", 0)
        del cache
    if "gemm" in what:
      for two in (0, 1, 2):
        lib.care_ctx_set_option(h, b"gemm_2sm", two)
        print("gemm_2sm =", two)
        for (M, N, Kd, name, odt) in [(R, 3 * d, d, "qkv", torch.bfloat16), (R, d, d, "out-proj", torch.float32),
                                      (R, 4 * d, d, "ffn1", torch.bfloat16), (R, d, 4 * d, "ffn2", torch.float32),
                                      (R, V, d, "vocab", torch.float32), (B * Lm, 2 * d, d, "cross-kv", torch.bfloat16)]:
            A = torch.randn(M, Kd, device="cuda").bfloat16()
            W = torch.randn(N, Kd, device="cuda").bfloat16()
            ldc = (N + 7) // 8 * 8
            C = torch.empty(M, ldc, device="cuda", dtype=odt)
            ms = timed(lambda: _lib.check(lib.care_gemm(
                h, BF16, A.data_ptr(), Kd, W.data_ptr(), Kd, None, C.data_ptr(), ldc,
                F32 if odt == torch.float32 else BF16, M, N, Kd, 0, st), "g"))
            err = (C[:512, :N].float() - A[:512].float() @ W.float().t()).abs().max().item()
            print("gemm %-9s M=%d N=%d K=%d: %.3f ms  %.0f TFLOP/s  (max err first 512 rows %.3g)" % (
                name, M, N, Kd, ms, 2.0 * M * N * Kd / ms / 1e9, err))
            del A, W, C
    if "beam" in what:
        from care_b200._lib import BeamState
        need = K
        ldv = (V + 7) // 8 * 8
        t = dict(
            scores=torch.zeros(B, K), cur_tok=torch.zeros(B * K, dtype=torch.int32),
            tok_hist=torch.zeros(B, Tm + 1, K, dtype=torch.int32), prev_ks=torch.zeros(B, Tm, K, dtype=torch.int32),
            anc=torch.zeros(B, K, Tm, dtype=torch.uint8), fin_score=torch.zeros(B, need),
            fin_t=torch.zeros(B, need, dtype=torch.int32), fin_k=torch.zeros(B, need, dtype=torch.int32),
            fin_count=torch.zeros(B, dtype=torch.int32), done=torch.zeros(B, dtype=torch.int32),
            n_done=torch.zeros(1, dtype=torch.int32), scratch=torch.zeros(B * K * 20))
        t = {k: v.cuda() for k, v in t.items()}
        bst = BeamState(B=B, K=K, T_max=Tm, V=V, need=need, **{k: v.data_ptr() for k, v in t.items()})
        logits = torch.randn(R, ldv, device="cuda")
        lib.care_beam_init(h, ctypes.byref(bst), 2, st)
        ms = timed(lambda: _lib.check(lib.care_beam_step(h, ctypes.byref(bst), logits.data_ptr(), ldv, 5, 30, None,
                                                         None, st), "b"))
        x = torch.randn(R, d, device="cuda").bfloat16()
        W = (torch.randn(V, d, device="cuda") * 0.05).bfloat16()
        nseg = lib.care_vocab_beam_nseg(h, R, V)
        part = torch.empty(R, nseg, 14, device="cuda")
        msf = timed(lambda: _lib.check(lib.care_vocab_beam_partials(h, x.data_ptr(), d, W.data_ptr(), d, R, V, d, K,
                                                                    part.data_ptr(), nseg, st), "f"))
        print("fused vocab+partials (nseg=%d): %.3f ms  %.0f TFLOP/s" % (nseg, msf, 2.0 * R * V * d / msf / 1e9))
        msu = timed(lambda: _lib.check(lib.care_beam_step_partials(h, ctypes.byref(bst), part.data_ptr(), nseg, 5, 30,
                                                                   None, None, st), "u"))
        print("beam update from partials: %.3f ms" % msu)
        nbytes = R * V * 4
        print("beam step: %.3f ms  %.0f GB/s (%.1f%%)" % (ms, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / PEAK))


if __name__ == "__main__":
    main()
