"""Host-side driver of the CUDA hot path: weight preparation, workspaces and the launch sequence.

PyTorch is plumbing here (device memory, streams); every arithmetic step is a call into
libcare_b200.so through the C ABI (care_b200/_lib.py).  There is no fallback path.

Precision modes
  "fp32": activations/weights fp32, SIMT FFMA GEMMs - the bit-exact token parity mode.
  "fp16": (default throughput mode) 16-bit tensor-core operands in IEEE fp16, fp32 accumulation; LayerNorm /
          softmax statistics, logits and beam scores stay fp32.  Every activation of this model is O(1) after a
          LayerNorm, so fp16's 10 mantissa bits cost no range and give 8x less rounding error than bf16;
          weights stored in bf16 are exactly representable.
  "bf16": the same kernels over bf16 (libcare_b200_bf16.so).
In both 16-bit modes everything upstream of a discrete ranking (encoder streams -> concept scores -> top-30
concepts; length predictor) runs as three-term split products on the tensor cores (care_split_f32_h16):
fp32-grade results, so the concept ids equal the fp32 mode's except on exact ties.
"""
import ctypes
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, BF16, F16, F32, BeamState, NextStep, check, ptr

PAD, UNK, BOS, EOS, MASK, VIS = 0, 1, 2, 3, 4, 5  # reference: config/Constants.py:1-6


def _round_up(x, m):
    return (x + m - 1) // m * m


class CareEngine:
    """One engine per (model, device, precision)."""

    def __init__(self, opt: dict, state_dict: Dict[str, torch.Tensor], device, precision: str = "fp16"):
        if precision not in ("fp32", "fp16", "bf16"):
            raise ValueError("precision must be 'fp32', 'fp16' or 'bf16'")
        self.lib = _lib.load("bf16" if precision == "bf16" else "fp16")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("care_b200 runs on CUDA devices only (no CPU path)")
        self.opt = dict(opt)
        self.precision = precision
        self.dt = {"fp32": F32, "fp16": F16, "bf16": BF16}[precision]
        self.tdtype = {"fp32": torch.float32, "fp16": torch.float16, "bf16": torch.bfloat16}[precision]
        self.half = precision != "fp32"
        # 16-bit modes: GEMMs upstream of a discrete ranking as 3-term split products (1 = plain 16-bit operands)
        self.enc_terms = int(opt.get("care_encoder_terms", 3)) if self.half else 1
        if self.enc_terms not in (1, 3):
            raise ValueError("care_encoder_terms must be 1 or 3")
        index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.device = torch.device("cuda", index)
        handle = ctypes.c_void_p()
        check(self.lib.care_ctx_create(ctypes.byref(handle), index), "care_ctx_create")
        self.ctx = handle
        if opt.get("care_self_compact") is not None:   # 0 / 1 / 2, see include/care_b200.h (library default: 2)
            check(self.lib.care_ctx_set_option(self.ctx, b"self_compact", int(opt["care_self_compact"])),
                  "care_ctx_set_option")
        self.d = opt["dim_hidden"]
        self.H = opt["num_attention_heads"]
        self.F = opt["intermediate_size"]
        self.V = opt["vocab_size"]
        self.eps = float(opt["layer_norm_eps"])
        self.modality = opt["modality"]
        self.m_dec = opt.get("modality_for_decoder") or self.modality
        self.m_pred = opt.get("modality_for_predictor") or self.modality
        self.max_len = opt.get("max_len", 30)
        self.ldv = _round_up(self.V, 8)
        # 16-bit modes: the vocabulary GEMM's epilogue feeds the beam kernel directly (no logits in HBM)
        self.fused_vocab = self.half and bool(opt.get("care_fused_vocab", True))
        # 16-bit modes: out-proj / FFN2 + residual + LayerNorm as one cluster kernel (care_gemm_add_ln):
        # 0 = care_gemm + care_add_ln, 1 = fused with a 16-bit residual stream, 2 = fused with an fp32 residual stream
        default_ln = int(os.environ.get("CARE_B200_FUSED_LN", "1"))
        fl = opt.get("care_fused_ln")
        self.fused_ln = int(default_ln if fl is None else fl) if (self.half and self.d in (512, 768, 1024)) else 0
        if self.fused_ln not in (0, 1, 2):
            raise ValueError("care_fused_ln must be 0, 1 or 2")
        self.fused_ln_min_rows = int(opt.get("care_fused_ln_min_rows", os.environ.get("CARE_B200_FUSED_LN_MIN_ROWS", 2048)))
        self._ws = {}
        self._ws_epoch_of = {}
        self._ws_bytes = 0
        self._epoch = 0            # advanced once per encoded batch (encode())
        limit = opt.get("care_workspace_limit_bytes")
        self._ws_limit = int(limit) if limit is not None else torch.cuda.get_device_properties(index).total_memory // 2
        self._nseg = {}
        # launch-bound regime (few rows per step): replay the whole decode as one CUDA graph
        self.use_graphs = bool(opt.get("care_cuda_graph", True))
        self.graph_max_rows = int(opt.get("care_cuda_graph_max_rows", 6144))
        # larger batches are graphed from the SECOND decode of the same shape on (a one-off shape is not worth a
        # capture): the replay has no launch gaps and no host polling (the device-side early-exit flag replaces
        # it) - 4096 videos: 56.2 -> 54.0 ms per batch
        self.graph_max_rows_repeat = int(opt.get("care_cuda_graph_max_rows_repeat",
                                                 os.environ.get("CARE_B200_GRAPH_REPEAT_ROWS", 32768)))
        self._shapes_seen = {}
        self._graphs = {}
        self._graph_launches = 0
        # graph-replayed decodes may run as concurrent lanes on separate streams (A/B switch, see _lanes_for);
        # every lane but the first owns a twin engine (own ctx + workspaces)
        self.graph_lanes = int(opt.get("care_graph_lanes", os.environ.get("CARE_B200_GRAPH_LANES", "1")))
        self._twins = []
        self._lane_streams = []
        # A/B switch: the beam kernel of step t also writes step t+1's decoder input rows (no embed_ln launch).  Off by
        # default - measured slower (4096 videos 51.2 -> 51.8 ms, 512 videos 8.90 -> 9.29 ms): one warp per video
        # embeds its K rows one after the other, where embed_ln_kernel has a warp per row
        self.fuse_next_step = bool(int(opt.get("care_fuse_next_step", os.environ.get("CARE_B200_FUSE_NEXT", "0"))))
        # live-slot records of the stream self-attention: 0 = a kernel of their own, 1 = written by the previous step's beam
        # kernel (measured slower), 2 = written by extra CTAs of the step's embedding launch (care_ctx_request_records)
        fi = opt.get("care_fuse_info")
        fi = int(os.environ.get("CARE_B200_FUSE_INFO", "2")) if fi is None else int(fi)
        check(self.lib.care_ctx_set_option(self.ctx, b"fuse_info", fi), "care_ctx_set_option")
        self.fuse_records = fi == 2
        self._x0_ready = None      # (t, R): x0 of step t was written by the previous step's beam kernel
        # step 1 of the beam search on one row per video (all K beams hold <bos>; 16-bit fused path)
        self.compact_first = bool(opt.get("care_compact_first_step", True))
        self._prepare_weights(state_dict)

    def __del__(self):
        try:
            if getattr(self, "ctx", None):
                self.lib.care_ctx_destroy(self.ctx)
                self.ctx = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    # weights
    # ------------------------------------------------------------------------------------------
    def _prepare_weights(self, sd):
        dev, T = self.device, self.tdtype
        opt, d = self.opt, self.d

        def f32(name):
            return sd[name].detach().to(dev, torch.float32).contiguous()

        def mat(t, pad_k=None):
            t = t.detach().to(dev, torch.float32)
            if pad_k is not None and t.shape[1] != pad_k:
                t = torch.nn.functional.pad(t, (0, pad_k - t.shape[1]))
            return t.to(T).contiguous()

        def head32(t, pad_k=None):
            """Weights of the small ranking heads (concept scores, GSG vector, length logits; M = videos): kept in
            fp32 and multiplied by the FFMA kernel in every mode - the tensor cores' truncating fp32 accumulation
            (~2^-18 over a few thousand k) is enough to swap neighbouring concepts, and these GEMMs are ~1 % of a
            batch."""
            t = t.detach().to(dev, torch.float32)
            if pad_k is not None and t.shape[1] != pad_k:
                t = torch.nn.functional.pad(t, (0, pad_k - t.shape[1]))
            return t.contiguous()

        def mat3(t):
            """Weight of a GEMM whose A operand comes from care_split_f32_h16(terms=3): [W_hi | W_hi | W_lo * 2^11],
            each part zero padded to a multiple of 64 columns (one swizzled TMA row)."""
            if self.enc_terms == 1:
                return mat(t, pad_k=_round_up(t.shape[1], 8) if self.half else None)
            t = t.detach().to(dev, torch.float32)
            kp = _round_up(t.shape[1], 64)
            t = torch.nn.functional.pad(t, (0, kp - t.shape[1]))
            hi = t.to(T)
            lo = ((t - hi.float()) * 2048.0).to(T)
            return torch.cat([hi, hi, lo], 1).contiguous()

        w = {}
        self.highway = opt["encoder"] == "EncoderWithHighWayBN"
        if opt["encoder"] not in ("Embedder", "EncoderWithHighWayBN"):
            raise ValueError("encoder %r is outside the accelerated hot path" % opt["encoder"])
        self.streams = []
        for ch in self.modality:
            p = "encoder.Encoder_%s" % ch.upper()
            s = dict(ch=ch, W=mat3(sd[p + ".0.weight"]), b=f32(p + ".0.bias"), dim=opt["dim_" + ch])
            if self.highway:
                s.update(W1=mat3(sd[p + ".1.w1.weight"]), b1=f32(p + ".1.w1.bias"),
                         W2=mat3(sd[p + ".1.w2.weight"]), b2=f32(p + ".1.w2.bias"),
                         bn_mean=f32(p + ".2.bn.running_mean"), bn_var=f32(p + ".2.bn.running_var"),
                         bn_w=f32(p + ".2.bn.weight"), bn_b=f32(p + ".2.bn.bias"))
            else:
                s.update(g=f32(p + ".1.weight"), beta=f32(p + ".1.bias"))
            self.streams.append(s)

        # predictor nets, ordered as the reference builds them (models/Predictor/__init__.py:26-60)
        nets = [c for c in opt["crits"] if c != "lang"] + list(opt.get("predictors_to_be_added", []))
        if opt.get("load_teacher_weights", False) and "length" in nets:
            nets.remove("length")
            nets.append("length")
        self.nets = nets
        self.n_attr = opt.get("attribute_prediction_k", 0)
        self.n_concepts = opt.get("use_attr_topk", 0) if "SemanticContainer" in nets else 0
        for i, kind in enumerate(nets):
            p = "predictor.nets.%d" % i
            if kind == "attribute":
                if not (opt.get("attribute_prediction_channel_concat") and opt.get("attribute_prediction_mean_pooling")):
                    raise ValueError("only the mean-pooling + channel-concat concept head is accelerated")
                w["prj_W"] = head32(sd[p + ".prj.weight"])
                w["prj_b"] = f32(p + ".prj.bias")
            elif kind == "SemanticContainer":
                w["attr_word"] = f32(p + ".attr_embs.word_embeddings.weight")
                w["attr_pos"] = f32(p + ".attr_embs.position_embeddings.weight")
                w["attr_g"] = f32(p + ".attr_embs.LayerNorm.weight")
                w["attr_b"] = f32(p + ".attr_embs.LayerNorm.bias")
                if "emb" in opt.get("use_attr_type", ""):
                    w["s2h_W"] = head32(sd[p + ".semantic2hidden.weight"], pad_k=_round_up(self.n_attr, 64))
            elif kind == "length":
                w["len_W0"] = head32(sd[p + ".net.0.weight"]); w["len_b0"] = f32(p + ".net.0.bias")
                w["len_W3"] = head32(sd[p + ".net.3.weight"]); w["len_b3"] = f32(p + ".net.3.bias")
            else:
                raise ValueError("predictor %r is outside the accelerated hot path" % kind)
        self.use_gsg = "s2h_W" in w
        self.concat_concepts = "concat" in opt.get("use_attr_type", "") and self.n_concepts > 0

        e = "decoder.embedding"
        w["word"] = f32(e + ".word_embeddings.weight")
        w["pos"] = f32(e + ".position_embeddings.weight")
        w["emb_g"] = f32(e + ".LayerNorm.weight")
        w["emb_b"] = f32(e + ".LayerNorm.bias")
        L = "decoder.layers.0."
        ia, xa = L + "intra_attention.", L + "inter_attention."
        w["Wqkv"] = mat(torch.cat([sd[ia + "SDPA.query.weight"], sd[ia + "SDPA.key.weight"],
                                   sd[ia + "SDPA.value.weight"]], 0))
        w["bqkv"] = torch.cat([f32(ia + "SDPA.query.bias"), f32(ia + "SDPA.key.bias"), f32(ia + "SDPA.value.bias")])
        w["Wo"] = mat(sd[ia + "dense.weight"]); w["bo"] = f32(ia + "dense.bias")
        w["ln1_g"] = f32(ia + "LayerNorm.weight"); w["ln1_b"] = f32(ia + "LayerNorm.bias")
        w["Wxq"] = mat(sd[xa + "SDPA.query.weight"]); w["bxq"] = f32(xa + "SDPA.query.bias")
        w["Wxkv"] = mat(torch.cat([sd[xa + "SDPA.key.weight"], sd[xa + "SDPA.value.weight"]], 0))
        w["bxkv"] = torch.cat([f32(xa + "SDPA.key.bias"), f32(xa + "SDPA.value.bias")])
        if (xa + "LayerNorm.weight") in sd:   # absent with attr_layer_pos 'parallel' (Layers.py:107-108)
            w["Wxo"] = mat(sd[xa + "dense.weight"]); w["bxo"] = f32(xa + "dense.bias")
            w["ln2_g"] = f32(xa + "LayerNorm.weight"); w["ln2_b"] = f32(xa + "LayerNorm.bias")
        hb = xa + "SDPA.hybrid_bias"
        w["hybrid_bias"] = f32(hb) if hb in sd else None
        # attr_attention (CABase): a second cross-attention over the concept embeddings (Layers.py:117-119)
        aa = L + "attr_attention."
        self.attr_pos = None
        if (aa + "SDPA.query.weight") in sd:
            self.attr_pos = opt.get("attr_layer_pos", "cross2attr")
            if self.attr_pos not in ("cross2attr", "attr2cross", "parallel"):
                raise ValueError("attr_layer_pos %r is outside the accelerated hot path" % self.attr_pos)
            w["Waq"] = mat(sd[aa + "SDPA.query.weight"]); w["baq"] = f32(aa + "SDPA.query.bias")
            w["Wakv"] = mat(torch.cat([sd[aa + "SDPA.key.weight"], sd[aa + "SDPA.value.weight"]], 0))
            w["bakv"] = torch.cat([f32(aa + "SDPA.key.bias"), f32(aa + "SDPA.value.bias")])
            if self.attr_pos == "parallel":
                # Layers.py:188-201: LN(x + dense_x(ctx_video) + dense_a(ctx_concepts)) - the two output projections
                # as ONE GEMM over the concatenated contexts, K = 2d, with the layer's own LayerNorm as its tail
                w["Wpo"] = mat(torch.cat([sd[xa + "dense.weight"], sd[aa + "dense.weight"]], 1))
                w["bpo"] = f32(xa + "dense.bias") + f32(aa + "dense.bias")
                w["lnp_g"] = f32(L + "LayerNorm.weight"); w["lnp_b"] = f32(L + "LayerNorm.bias")
            else:
                w["Wao"] = mat(sd[aa + "dense.weight"]); w["bao"] = f32(aa + "dense.bias")
                w["lna_g"] = f32(aa + "LayerNorm.weight"); w["lna_b"] = f32(aa + "LayerNorm.bias")
        w["W1"] = mat(sd[L + "ffn.dense1.weight"]); w["b1"] = f32(L + "ffn.dense1.bias")
        w["W2"] = mat(sd[L + "ffn.dense2.weight"]); w["b2"] = f32(L + "ffn.dense2.bias")
        w["ln3_g"] = f32(L + "ffn.LayerNorm.weight"); w["ln3_b"] = f32(L + "ffn.LayerNorm.bias")
        w["Wvocab"] = mat(sd["cls_head.tgt_word_prj.weight"])
        self.w = w

        # memory layout: decoder modalities in modality order, then the concept embeddings
        self.frames = {ch: (opt.get("retrieval_topk", 20) if ch == "r" else opt["n_frames"]) for ch in self.modality}
        off = 0
        self.mem_off = {}
        for ch in self.modality:
            if ch in self.m_dec:
                self.mem_off[ch] = off
                off += self.frames[ch]
        self.enc_len = off
        self.Lm = off + (self.n_concepts if self.concat_concepts else 0)
        if w["hybrid_bias"] is not None and w["hybrid_bias"].shape[1] != self.Lm:
            raise ValueError("hybrid_bias length %d != memory length %d" % (w["hybrid_bias"].shape[1], self.Lm))

    # ------------------------------------------------------------------------------------------
    # helpers
    # ------------------------------------------------------------------------------------------
    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _buf(self, name, shape, dtype, zero=False):
        """Workspace keyed by (name, shape, dtype), kept across calls.  A service that sees many batch
        sizes would otherwise accumulate one workspace set per size: once the sets exceed
        `care_workspace_limit_bytes` (default: half of the device memory), the buffers the current
        call has not touched are dropped (stream-ordered reuse makes that safe for launches in flight)."""
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            nbytes = torch.empty((), dtype=dtype).element_size()
            for n in shape:
                nbytes *= int(n)
            if self._ws_bytes + nbytes > self._ws_limit:
                self._evict_stale()
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self._ws[key] = t
            self._ws_bytes += nbytes
        self._ws_epoch_of[key] = self._epoch
        return t

    def _evict_stale(self):
        stale = [k for k, e in self._ws_epoch_of.items() if e < self._epoch]
        if not stale:
            return
        (getattr(self, "_owner", None) or self)._graphs.clear()   # graphs hold raw pointers into the workspaces
        for k in stale:
            t = self._ws.pop(k)
            self._ws_bytes -= t.numel() * t.element_size()
            del self._ws_epoch_of[k]

    def free_workspaces(self):
        self._graphs.clear()   # graphs hold raw pointers into the workspaces
        self._ws.clear()
        self._ws_epoch_of.clear()
        self._ws_bytes = 0
        for t in self._twins:
            t.free_workspaces()

    def copy_stream(self):
        """Side stream for host->device feature copies that overlap the decode."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        return self._copy_stream

    def launch_count(self):
        """Kernels launched so far: eager launches counted by the library + launches replayed by graphs."""
        eager = sum(int(self.lib.care_ctx_launch_count(e.ctx)) for e in [self] + self._twins)
        return eager + self._graph_launches

    def gemm(self, A, W, bias, C, M, N, K, act=ACT_NONE, lda=None, ldc=None, dt=None):
        out_dt = F32 if C.dtype == torch.float32 else self.dt
        check(self.lib.care_gemm(self.ctx, self.dt if dt is None else dt, ptr(A), lda if lda is not None else A.stride(-2), ptr(W),
                                 W.stride(0), ptr(bias), ptr(C), ldc if ldc is not None else C.stride(-2), out_dt,
                                 M, N, K, act, self._stream()), "care_gemm")

    def _fused_for(self, M):
        return self.fused_ln == 2 or (self.fused_ln == 1 and M >= self.fused_ln_min_rows)

    def _sublayer_tail(self, A, W, bias, gamma, beta, res, out, M, K, y32, res32=None, out32=None):
        """out = LayerNorm(A W^T + bias + res) (SubLayers.py:68-79,137-152): one cluster kernel in the 16-bit
        modes, care_gemm -> fp32 y32 -> care_add_ln otherwise.  res32 / out32: the fp32 residual stream
        (care_fused_ln = 2)."""
        d = self.d
        # the cluster kernel needs a couple of thousand rows to fill the SMs (128-row blocks x d/256 CTAs): below
        # that the unfused pair with its narrower tiles (or the weight-streaming kernel for a handful of rows) is
        # faster (measured: 64 videos 5.68 vs 5.89 ms, one video 3.7 vs 4.3 ms per caption)
        if self._fused_for(M):
            r32 = self.fused_ln == 2
            check(self.lib.care_gemm_add_ln(self.ctx, ptr(A), A.stride(-2), ptr(W), W.stride(0), ptr(bias),
                                            ptr(res32 if r32 else res), F32 if r32 else self.dt, ptr(gamma), ptr(beta),
                                            self.eps, ptr(out), ptr(out32) if r32 else None, M, d, K, self._stream()),
                  "care_gemm_add_ln")
            return
        self.gemm(A, W, bias, y32, M, d, K)
        check(self.lib.care_add_ln(self.ctx, self.dt, ptr(y32), ptr(res), ptr(gamma), ptr(beta), self.eps, M, d, ptr(out),
                                   self._stream()), "care_add_ln")

    def _operand(self, name, x32, rows, cols):
        """fp32 [rows, cols] -> the A operand of an upstream-of-ranking GEMM and its K: fp32 mode: x itself;
        16-bit modes: [hi | lo | hi] (or a plain cast when care_encoder_terms == 1)."""
        if not self.half:
            return x32, cols
        kp = _round_up(cols, 64) if self.enc_terms == 3 else _round_up(cols, 8)
        a = self._buf(name, (rows, self.enc_terms * kp), self.tdtype)
        check(self.lib.care_split_f32_h16(self.ctx, ptr(x32), x32.stride(-2), rows, cols, kp, self.enc_terms, ptr(a),
                                          self._stream()), "care_split_f32_h16")
        return a, self.enc_terms * kp

    # ------------------------------------------------------------------------------------------
    # encoding phase  (reference: models/Framework.py:150-187)
    # ------------------------------------------------------------------------------------------
    def encode(self, feats: List[torch.Tensor]) -> Dict[str, torch.Tensor]:
        opt, d, T, w = self.opt, self.d, self.tdtype, self.w
        feats = feats[:len(self.modality)]
        B = feats[0].shape[0]
        self._epoch += 1
        st = self._stream()
        lib, ctx, dt = self.lib, self.ctx, self.dt
        n_pred = len([c for c in self.modality if c in self.m_pred])
        has_attr = "attribute" in self.nets
        memory = torch.empty((B, self.Lm, d), dtype=T, device=self.device)
        means = (self._buf("means", (B, max(n_pred, 1) * d), torch.float32)
                 if has_attr or "length" in self.nets else None)
        pred_slot = 0
        pred_tokens = []
        for s, x in zip(self.streams, feats):
            ch, dim = s["ch"], s["dim"]
            if x.dim() != 3 or x.shape[0] != B or x.shape[1] != self.frames[ch] or x.shape[2] != dim:
                # the reference fails inside nn.Linear / torch.cat on such a batch; here the rows of the memory
                # buffer are laid out from the opt, so a mismatch must not reach the kernels
                raise ValueError("feats[%r] has shape %s, expected [%d, %d, %d] (n_frames / retrieval_topk / dim_%s "
                                 "of the checkpoint's opt)" % (ch, tuple(x.shape), B, self.frames[ch], dim, ch))
            Tn = x.shape[1]
            if x.dtype != torch.float32 or not x.is_contiguous():
                x = x.to(torch.float32).contiguous()
            rows = B * Tn
            a, ka = self._operand("feat_h16", x.view(rows, dim), rows, dim)
            h32 = self._buf("enc_h32", (rows, d), torch.float32)
            self.gemm(a, s["W"], s["b"], h32, rows, d, ka)
            in_dec, in_pred = ch in self.m_dec, ch in self.m_pred and means is not None
            out_ptr = ptr(memory) if in_dec else None
            mean_ptr = ptr(means) if in_pred else None
            if self.highway:
                hT, kh = self._operand("enc_hT", h32, rows, d)
                y32 = self._buf("enc_y32", (rows, d), torch.float32)
                g32 = self._buf("enc_g32", (rows, d), torch.float32)
                self.gemm(hT, s["W1"], s["b1"], y32, rows, d, kh)
                self.gemm(hT, s["W2"], s["b2"], g32, rows, d, kh)
                check(lib.care_encoder_highway_bn_mean(
                    ctx, dt, ptr(h32), ptr(y32), ptr(g32), ptr(s["bn_mean"]), ptr(s["bn_var"]), ptr(s["bn_w"]),
                    ptr(s["bn_b"]), 1e-5, B, Tn, d, out_ptr, self.Lm, self.mem_off.get(ch, 0), mean_ptr,
                    n_pred * d, pred_slot * d, st), "care_encoder_highway_bn_mean")
            else:
                check(lib.care_encoder_ln_mean(
                    ctx, dt, ptr(h32), ptr(s["g"]), ptr(s["beta"]), self.eps, B, Tn, d, out_ptr, self.Lm,
                    self.mem_off.get(ch, 0), mean_ptr, n_pred * d, pred_slot * d, st), "care_encoder_ln_mean")
            if in_pred:
                pred_slot += 1
                pred_tokens.append(Tn)
        out = {"encoder_hidden_states": memory}
        if has_attr:
            ld_sc = _round_up(self.n_attr, 8)
            scores = self._buf("attr_scores", (B, ld_sc), torch.float32)
            self.gemm(means, w["prj_W"], w["prj_b"], scores, B, self.n_attr, n_pred * d, dt=F32)
            preds = torch.empty((B, self.n_attr), dtype=torch.float32, device=self.device)
            out["preds_attr"] = preds
            if "SemanticContainer" in self.nets:
                predsT = kpad = None
                if self.use_gsg:
                    kpad = w["s2h_W"].shape[1]
                    predsT = self._buf("preds_T", (B, kpad), torch.float32)
                labels = torch.empty((B, self.n_concepts), dtype=torch.int64, device=self.device)
                sem = None
                if not self.concat_concepts and self.attr_pos is not None:
                    sem = torch.empty((B, self.n_concepts, d), dtype=T, device=self.device)
                    out["semantic_embs"] = sem
                check(lib.care_concept_head(
                    ctx, dt, ptr(scores), ld_sc, B, self.n_attr, self.n_concepts, ptr(w["attr_word"]),
                    ptr(w["attr_pos"]), ptr(w["attr_g"]), ptr(w["attr_b"]), self.eps, d, ptr(preds), ptr(predsT),
                    kpad or 0, ptr(labels), ptr(memory) if self.concat_concepts else ptr(sem),
                    self.Lm if self.concat_concepts else self.n_concepts, self.enc_len if self.concat_concepts else 0,
                    st), "care_concept_head")
                out["semantic_labels"] = labels
                if self.use_gsg:
                    gsg = torch.empty((B, d), dtype=torch.float32, device=self.device)
                    self.gemm(predsT, w["s2h_W"], None, gsg, B, d, kpad, dt=F32)
                    out["semantic_hidden_states"] = gsg
            else:
                check(lib.care_concept_head(
                    ctx, dt, ptr(scores), ld_sc, B, self.n_attr, 1, None, None, None, None, self.eps, d, ptr(preds),
                    None, 0, None, None, 0, 0, st), "care_concept_head")
        if "length" in self.nets:
            # Predictor_length (pred_length.py:14-22): mean over all predictor tokens, then a 2-layer MLP.
            out["preds_length_logits"] = self._length_logits(B, means, pred_tokens, n_pred)
        return out

    def _length_logits(self, B, means, pred_tokens, n_pred):
        """Predictor_length (pred_length.py:5-22): Linear -> ReLU -> Linear over the mean of all predictor
        tokens, obtained from the per-stream means weighted by their token counts.  Returns fp32 logits
        [B, >= max_len] (log_softmax is monotone; only the ranking is consumed, Translator.py:307-311)."""
        w, d = self.w, self.d
        total = float(sum(pred_tokens))
        weights = (ctypes.c_float * len(pred_tokens))(*[t / total for t in pred_tokens])
        comb = self._buf("len_in", (B, d), torch.float32)
        check(self.lib.care_combine_means(self.ctx, F32, ptr(means), B, n_pred, d, weights, ptr(comb), d,
                                          self._stream()), "care_combine_means")
        hid = self._buf("len_hid", (B, d), torch.float32)
        self.gemm(comb, w["len_W0"], w["len_b0"], hid, B, d, d, act=ACT_RELU, dt=F32)
        logits = torch.empty((B, _round_up(self.max_len, 8)), dtype=torch.float32, device=self.device)
        self.gemm(hid, w["len_W3"], w["len_b3"], logits, B, self.max_len, d, dt=F32)
        return logits

    def cross_kv(self, memory, static=False):
        """K/V of the cross-attention memory, projected once per video (hoisted out of the step loop)."""
        B = memory.shape[0]
        if static:
            kv = self._buf("cross_kv", (B, self.Lm, 2 * self.d), self.tdtype)
        else:
            kv = torch.empty((B, self.Lm, 2 * self.d), dtype=self.tdtype, device=self.device)
        self.gemm(memory.view(B * self.Lm, self.d), self.w["Wxkv"], self.w["bxkv"], kv.view(B * self.Lm, 2 * self.d),
                  B * self.Lm, 2 * self.d, self.d)
        return kv

    def attr_kv(self, enc, static=False):
        """K/V of the concept embeddings for the attr_attention block, projected once per video."""
        if self.attr_pos is None:
            return None
        sem = enc.get("semantic_embs")
        if sem is None:
            raise KeyError("this model has an attr_attention layer: `semantic_embs` is required")
        if sem.dtype != self.tdtype or not sem.is_contiguous():
            sem = sem.to(self.tdtype).contiguous()
        B, n = sem.shape[0], sem.shape[1]
        if static:
            akv = self._buf("attr_kv", (B, n, 2 * self.d), self.tdtype)
        else:
            akv = torch.empty((B, n, 2 * self.d), dtype=self.tdtype, device=self.device)
        self.gemm(sem.view(B * n, self.d), self.w["Wakv"], self.w["bakv"], akv.view(B * n, 2 * self.d), B * n,
                  2 * self.d, self.d)
        return akv

    def _attr_block_step(self, x_in, x_out, akv, B, K, done, x_in32=None, x_out32=None):
        """attr_attention for the newest position of every beam row: LN(dense(attn(q(x), concepts)) + x)."""
        lib, ctx, dt, w, d, T = self.lib, self.ctx, self.dt, self.w, self.d, self.tdtype
        st = self._stream()
        R = B * K
        qa = self._buf("qa", (R, d), T); cxa = self._buf("ctx_a", (R, d), T)
        y32 = None if self._fused_for(R) else self._buf("y32", (R, d), torch.float32)
        self.gemm(x_in, w["Waq"], w["baq"], qa, R, d, d)
        check(lib.care_cross_attn_step(ctx, dt, ptr(qa), d, ptr(akv), akv.shape[1], B, K, self.H, d, None, done,
                                       ptr(cxa), st), "care_cross_attn_step(attr)")
        self._sublayer_tail(cxa, w["Wao"], w["bao"], w["lna_g"], w["lna_b"], x_in, x_out, R, d, y32, x_in32, x_out32)

    # ------------------------------------------------------------------------------------------
    # auto-regressive beam decode  (reference: models/Translator.py:35-220 + misc/Decoding/Beam.py)
    # ------------------------------------------------------------------------------------------
    def _beam_buffers(self, B, K, need):
        Tm = self.max_len - 1
        dev = self.device
        i32 = torch.int32
        bufs = dict(
            scores=self._buf("bs_scores", (B, K), torch.float32),
            cur_tok=self._buf("bs_cur", (B * K,), i32),
            tok_hist=self._buf("bs_hist", (B, Tm + 1, K), i32),
            prev_ks=self._buf("bs_prev", (B, Tm, K), i32),
            anc=self._buf("bs_anc", (B, K, Tm), torch.uint8),
            fin_score=self._buf("bs_fs", (B, need), torch.float32),
            fin_t=self._buf("bs_ft", (B, need), i32),
            fin_k=self._buf("bs_fk", (B, need), i32),
            fin_count=self._buf("bs_fc", (B,), i32),
            done=self._buf("bs_done", (B,), i32),
            n_done=self._buf("bs_nd", (1,), i32),
            scratch=self._buf("bs_scratch", (B * K * 20,), torch.float32),
        )
        st = BeamState(B=B, K=K, T_max=Tm, V=self.V, need=need, **{k: ptr(v) for k, v in bufs.items()})
        return bufs, st

    def step_hidden(self, t, B, K, enc, kv, bufs, akv=None, compact_first=False):
        """Decoder layer for the newest position of every beam row (10 launches with the fused residual LayerNorm, 13
        without); returns the hidden states [R, d] the vocabulary projection consumes.  `bufs` holds the shared beam
        state (tokens, ancestry).
        compact_first (t == 1 only): before the first step all K beams of a video hold <bos> and identical state, and
        Beam.advance looks at beam 0 only (Beam.py:56): the layer runs on ONE row per video ([B, d] result).  Its
        q|k|v land in cache slot 0 of position 0 - the only slot of that position any later step references, because
        every beam chosen at step 1 descends from beam 0 - and its self-attention over the single <bos> key is the
        value row itself (a softmax over one key is exactly 1)."""
        lib, ctx, dt, w, d, T = self.lib, self.ctx, self.dt, self.w, self.d, self.tdtype
        st = self._stream()
        Tm = self.max_len - 1
        # zero-filled when first allocated: the compact first step writes slot 0 of position 0 only, and the dense
        # self-attention tile loads every slot of a position (masked keys weigh exactly 0, but 0 x NaN garbage = NaN);
        # afterwards the other slots hold finite values of an earlier decode at worst
        cache = self._buf("kv_cache", (Tm, B * K, 3 * d), T, zero=True)
        if compact_first:
            assert t == 1
            K_state, K = K, 1
        R = B * K
        x0 = self._buf("x0", (R, d), T); x1 = self._buf("x1", (R, d), T)
        x2 = self._buf("x2", (R, d), T); x3 = self._buf("x3", (R, d), T)
        cx = self._buf("ctx", (R, d), T); qc = self._buf("qc", (R, d), T)
        y32 = None if self._fused_for(R) else self._buf("y32", (R, d), torch.float32)
        hb = self._buf("ffn_h", (R, self.F), T)
        # fp32 residual stream (care_fused_ln = 2): r0..r3 next to the 16-bit GEMM operands x0..x3
        r0 = r1 = r2 = r3 = ra = None
        if self.fused_ln == 2:
            r0 = self._buf("r0", (R, d), torch.float32); r1 = self._buf("r1", (R, d), torch.float32)
            r2 = self._buf("r2", (R, d), torch.float32); r3 = self._buf("r3", (R, d), torch.float32)
        gsg = enc.get("semantic_hidden_states")
        done = ptr(bufs["done"])
        if compact_first:
            tok0 = self._buf("first_tok", (B,), torch.int32)
            tok0.copy_(bufs["cur_tok"].view(B, K_state)[:, 0])
            slot0 = cache[0].view(B, K_state, 3 * d)[:, 0, :]          # [B, 3d], row stride K * 3d
            check(lib.care_embed_ln(ctx, dt, ptr(tok0), None, 0, ptr(w["word"]), ptr(w["pos"]), None, ptr(gsg), 1,
                                    ptr(w["emb_g"]), ptr(w["emb_b"]), self.eps, R, d, ptr(x0), ptr(r0), st),
                  "care_embed_ln")
            self.gemm(x0, w["Wqkv"], w["bqkv"], slot0, R, 3 * d, d)
            cx_in = slot0[:, 2 * d:]                                   # the value rows: attention over one key
        else:
            if self._x0_ready == (t, R):
                self._x0_ready = None      # written by the beam kernel of step t - 1 (care_ctx_set_next_step)
            else:
                if self.fuse_records:
                    check(lib.care_ctx_request_records(ctx, ptr(bufs["anc"]), Tm, ptr(bufs["tok_hist"]), done, B, K, self.H, t),
                          "care_ctx_request_records")
                check(lib.care_embed_ln(ctx, dt, ptr(bufs["cur_tok"]), None, t - 1, ptr(w["word"]), ptr(w["pos"]), None,
                                        ptr(gsg), K, ptr(w["emb_g"]), ptr(w["emb_b"]), self.eps, R, d, ptr(x0), ptr(r0),
                                        st), "care_embed_ln")
            self.gemm(x0, w["Wqkv"], w["bqkv"], cache[t - 1], R, 3 * d, d)
            check(lib.care_self_attn_step(ctx, dt, ptr(cache), t, B, K, self.H, d, ptr(bufs["anc"]), Tm,
                                          ptr(bufs["tok_hist"]), done, ptr(cx), st), "care_self_attn_step")
            cx_in = cx
        self._sublayer_tail(cx_in, w["Wo"], w["bo"], w["ln1_g"], w["ln1_b"], x0, x1, R, d, y32, r0, r1)
        if self.attr_pos == "attr2cross":   # Layers.py:180-187
            xa = self._buf("xa", (R, d), T)
            ra = self._buf("ra", (R, d), torch.float32) if self.fused_ln == 2 else None
            self._attr_block_step(x1, xa, akv, B, K, done, r1, ra)
            x1, r1 = xa, ra
        self.gemm(x1, w["Wxq"], w["bxq"], qc, R, d, d)
        check(lib.care_cross_attn_step(ctx, dt, ptr(qc), d, ptr(kv), self.Lm, B, K, self.H, d, ptr(w["hybrid_bias"]),
                                       done, ptr(cx), st), "care_cross_attn_step")
        if self.attr_pos == "parallel":   # Layers.py:188-201
            qa = self._buf("qa", (R, d), T); cxa = self._buf("ctx_a", (R, d), T)
            cat = self._buf("ctx_cat", (R, 2 * d), T)
            self.gemm(x1, w["Waq"], w["baq"], qa, R, d, d)
            check(lib.care_cross_attn_step(ctx, dt, ptr(qa), d, ptr(akv), akv.shape[1], B, K, self.H, d, None, done,
                                           ptr(cxa), st), "care_cross_attn_step(attr)")
            torch.cat([cx, cxa], dim=1, out=cat)
            self._sublayer_tail(cat, w["Wpo"], w["bpo"], w["lnp_g"], w["lnp_b"], x1, x2, R, 2 * d, y32, r1, r2)
        else:
            self._sublayer_tail(cx, w["Wxo"], w["bxo"], w["ln2_g"], w["ln2_b"], x1, x2, R, d, y32, r1, r2)
        if self.attr_pos == "cross2attr":   # Layers.py:217-225
            xa = self._buf("xa", (R, d), T)
            ra = self._buf("ra", (R, d), torch.float32) if self.fused_ln == 2 else None
            self._attr_block_step(x2, xa, akv, B, K, done, r2, ra)
            x2, r2 = xa, ra
        self.gemm(x2, w["W1"], w["b1"], hb, R, self.F, d, act=ACT_RELU)
        self._sublayer_tail(hb, w["W2"], w["b2"], w["ln3_g"], w["ln3_b"], x2, x3, R, self.F, y32, r2, r3)
        return x3

    def _arm_next_step(self, t, B, K, enc):
        """Asks the beam kernel of step t for step t+1's decoder input rows (x0 of the full K-row layout)."""
        if not self.fuse_next_step or t + 1 >= self.max_len:
            return
        w, d, R = self.w, self.d, B * K
        x0 = self._buf("x0", (R, d), self.tdtype)
        r0 = self._buf("r0", (R, d), torch.float32) if self.fused_ln == 2 else None
        nxt = NextStep(ptr(w["word"]), ptr(w["pos"]), ptr(enc.get("semantic_hidden_states")), ptr(w["emb_g"]),
                       ptr(w["emb_b"]), self.eps, d, ptr(x0), ptr(r0))
        check(self.lib.care_ctx_set_next_step(self.ctx, ctypes.byref(nxt)), "care_ctx_set_next_step")
        self._x0_ready = (t + 1, R)

    def decode_step(self, t, B, K, enc, kv, bufs, bst, audit=None, want_logits=False, akv=None):
        """One beam step (len_input_ids == t) for every video; 12 kernel launches from 2048 beam rows up (fused residual LayerNorm + fused vocabulary), 15 for small batches."""
        lib, ctx, w, d = self.lib, self.ctx, self.w, self.d
        st = self._stream()
        R = B * K
        if t == 1:
            self._x0_ready = None     # a decode that stopped early may have left a request behind
        fused = self.fused_vocab and not want_logits
        if t == 1 and fused and audit is None and K > 1 and self.compact_first:
            # step 1 on one row per video (see step_hidden): the vocabulary kernel and the beam kernel take B rows
            x3 = self.step_hidden(t, B, K, enc, kv, bufs, akv, compact_first=True)
            nseg = self._nseg.get(B)
            if nseg is None:
                nseg = self._nseg[B] = int(lib.care_vocab_beam_nseg(ctx, B, self.V))
            kb = 2 if K <= 1 else 4 if K <= 3 else 6 if K <= 5 else 9
            part = self._buf("vocab_partials", (B, nseg, 2 + 2 * kb), torch.float32)
            check(lib.care_vocab_beam_partials(ctx, ptr(x3), d, ptr(w["Wvocab"]), w["Wvocab"].stride(0), B, self.V, d,
                                               K, ptr(part), nseg, st), "care_vocab_beam_partials")
            self._arm_next_step(t, B, K, enc)
            check(lib.care_beam_first_step_partials(ctx, ctypes.byref(bst), ptr(part), nseg, self.max_len, None, None,
                                                    st), "care_beam_first_step_partials")
            return None
        x3 = self.step_hidden(t, B, K, enc, kv, bufs, akv)
        logits = None if fused else self._buf("logits", (R, self.ldv), torch.float32)
        cv = ci = None
        if audit is not None:
            cv, ci = audit
        if fused:
            nseg = self._nseg.get(R)
            if nseg is None:
                nseg = self._nseg[R] = int(lib.care_vocab_beam_nseg(ctx, R, self.V))
            kb = 2 if K <= 1 else 4 if K <= 3 else 6 if K <= 5 else 9
            part = self._buf("vocab_partials", (R, nseg, 2 + 2 * kb), torch.float32)
            check(lib.care_vocab_beam_partials(ctx, ptr(x3), d, ptr(w["Wvocab"]), w["Wvocab"].stride(0), R, self.V, d,
                                               K, ptr(part), nseg, st), "care_vocab_beam_partials")
            if audit is None:
                self._arm_next_step(t, B, K, enc)
            check(lib.care_beam_step_partials(ctx, ctypes.byref(bst), ptr(part), nseg, t, self.max_len, ptr(cv),
                                              ptr(ci), st), "care_beam_step_partials")
            return None
        self.gemm(x3, w["Wvocab"], None, logits, R, self.V, d)
        check(lib.care_beam_step(ctx, ctypes.byref(bst), ptr(logits), self.ldv, t, self.max_len, ptr(cv), ptr(ci),
                                 st), "care_beam_step")
        return logits

    def ar_decode(self, enc, B, beam_size=5, topk=1, beam_alpha=1.0, early_exit_every=4, trace=None,
                  trace_logits=False, bos=None):
        """Runs the whole beam search on device; returns padded int32 ids, lengths, raw scores, steps
        (device tensors).  `trace` (a list, tests only) receives per-step snapshots of the beam state."""
        K = beam_size
        need = max(K, topk)
        lib, ctx = self.lib, self.ctx
        bos = BOS if bos is None else int(bos)
        if trace is None and self.use_graphs:
            shape = (B, K, topk, float(beam_alpha), bos)
            seen = self._shapes_seen.get(shape, 0)
            self._shapes_seen[shape] = seen + 1
            if B * K <= self.graph_max_rows or (seen >= 1 and B * K <= self.graph_max_rows_repeat):
                return self._ar_decode_graph(enc, B, K, topk, float(beam_alpha), bos)
        st = self._stream()
        kv = self.cross_kv(enc["encoder_hidden_states"])
        akv = self.attr_kv(enc)
        bufs, bst = self._beam_buffers(B, K, need)
        check(lib.care_beam_init(ctx, ctypes.byref(bst), bos, st), "care_beam_init")
        lib.care_ctx_set_early_exit(ctx, ptr(bufs["n_done"]), B)
        try:
            for t in range(1, self.max_len):
                audit = None
                if trace is not None:
                    audit = (torch.empty((B, K + 1), dtype=torch.float32, device=self.device),
                             torch.empty((B, K + 1), dtype=torch.int32, device=self.device))
                    pre = {k: bufs[k].cpu().clone() for k in ("anc", "tok_hist", "done", "scores", "cur_tok")}
                logits = self.decode_step(t, B, K, enc, kv, bufs, bst, audit, want_logits=trace_logits, akv=akv)
                if trace is not None:
                    trace.append(dict(step=t, pre=pre, cand_val=audit[0].cpu(), cand_idx=audit[1].cpu(),
                                      logits=logits[:, :self.V].cpu().clone() if trace_logits else None))
                if early_exit_every and t % early_exit_every == 0 and t < self.max_len - 1:
                    if int(bufs["n_done"].item()) == B:
                        break
        finally:
            lib.care_ctx_set_early_exit(ctx, None, 0)
        Tm = self.max_len - 1
        out_tok = torch.empty((B, topk, Tm), dtype=torch.int32, device=self.device)
        out_len = torch.empty((B, topk), dtype=torch.int32, device=self.device)
        out_score = torch.empty((B, topk), dtype=torch.float32, device=self.device)
        out_t = torch.empty((B, topk), dtype=torch.int32, device=self.device)
        check(lib.care_beam_finalize(ctx, ctypes.byref(bst), float(beam_alpha), topk, ptr(out_tok), ptr(out_len),
                                     ptr(out_score), ptr(out_t), st), "care_beam_finalize")
        return out_tok, out_len, out_score, out_t

    def twin(self):
        """A second engine over the SAME weights with its own library context and workspaces: lets two halves of a
        small batch decode concurrently on two streams (see _ar_decode_graph)."""
        t = object.__new__(CareEngine)
        t.__dict__.update(self.__dict__)
        handle = ctypes.c_void_p()
        check(self.lib.care_ctx_create(ctypes.byref(handle), self.device.index), "care_ctx_create")
        t.ctx = handle
        check(self.lib.care_ctx_share_tuning(t.ctx, self.ctx), "care_ctx_share_tuning")
        t._owner = self
        if self.opt.get("care_self_compact") is not None:
            check(self.lib.care_ctx_set_option(t.ctx, b"self_compact", int(self.opt["care_self_compact"])),
                  "care_ctx_set_option")
        t._ws, t._ws_epoch_of, t._ws_bytes = {}, {}, 0
        t._graphs, t._graph_launches, t._twins, t._nseg = {}, 0, [], {}
        t._copy_stream = None
        return t

    def _lanes_for(self, B, K):
        """Concurrent lanes of a graph-replayed decode: independent slices of the batch on separate streams (option
        care_graph_lanes, default 1).  Measured on B200 (round 2, cfg4): 512 videos 9.35 ms with one lane, 9.84 / 11.2 /
        11.4 ms with 2 / 3 / 4; 256 videos 7.07 -> 7.84 ms; 64 videos unchanged - the persistent GEMM / vocabulary
        kernels of one lane already occupy every SM, so lanes serialise and only add their fixed costs.  Kept as an
        A/B switch."""
        return max(1, min(self.graph_lanes, B))

    def _ar_decode_graph(self, enc, B, K, topk, beam_alpha, bos=BOS):
        """The whole decode (cross K/V projection, beam init, max_len-1 steps, finalisation) as ONE CUDA
        graph per (B, K, topk, alpha): a fixed launch sequence over fixed workspaces with no host
        sync inside, so small batches are not bound by per-launch host overhead.  Inputs are copied
        into static buffers, outputs are cloned out of them.  Larger graph batches run as concurrent lanes
        (_lanes_for): videos are independent, so every lane is a complete decode of its slice of the batch."""
        Tm = self.max_len - 1
        need = max(K, topk)
        n_lanes = self._lanes_for(B, K)
        while len(self._twins) < n_lanes - 1:
            self._twins.append(self.twin())
        engines = [self] + self._twins[:n_lanes - 1]
        bounds = [(i * B // n_lanes, (i + 1) * B // n_lanes) for i in range(n_lanes)]
        if n_lanes > 1 and len(self._lane_streams) < n_lanes - 1:
            self._lane_streams += [torch.cuda.Stream(self.device) for _ in range(n_lanes - 1 - len(self._lane_streams))]
        lanes = []
        for eng, (lo, hi) in zip(engines, bounds):
            n = hi - lo
            eng._epoch = self._epoch
            memory = eng._buf("g_memory", (n, self.Lm, self.d), self.tdtype)
            memory.copy_(enc["encoder_hidden_states"][lo:hi])
            static_enc = {"encoder_hidden_states": memory, "semantic_hidden_states": None}
            if self.use_gsg and enc.get("semantic_hidden_states") is not None:
                gsg = eng._buf("g_gsg", (n, self.d), torch.float32)
                gsg.copy_(enc["semantic_hidden_states"][lo:hi])
                static_enc["semantic_hidden_states"] = gsg
            if self.attr_pos is not None:
                sem = eng._buf("g_sem", (n, self.n_concepts, self.d), self.tdtype)
                sem.copy_(enc["semantic_embs"][lo:hi])
                static_enc["semantic_embs"] = sem
            outs = (eng._buf("g_out_tok", (n, topk, Tm), torch.int32), eng._buf("g_out_len", (n, topk), torch.int32),
                    eng._buf("g_out_score", (n, topk), torch.float32), eng._buf("g_out_t", (n, topk), torch.int32))
            lanes.append(dict(eng=eng, n=n, enc=static_enc, outs=outs))

        def body():
            # the current stream is the capturing stream while the graph is recorded: the lanes fork from it
            main = torch.cuda.current_stream(self.device)
            streams = [main] + self._lane_streams[:n_lanes - 1]
            for side in streams[1:]:
                side.wait_stream(main)
            for lane, stream in zip(lanes, streams):
                eng, n = lane["eng"], lane["n"]
                with torch.cuda.stream(stream):
                    lane["kv"] = eng.cross_kv(lane["enc"]["encoder_hidden_states"], static=True)
                    lane["akv"] = eng.attr_kv(lane["enc"], static=True)
                    lane["bufs"], lane["bst"] = eng._beam_buffers(n, K, need)
                    check(eng.lib.care_beam_init(eng.ctx, ctypes.byref(lane["bst"]), bos, eng._stream()), "care_beam_init")
                    eng.lib.care_ctx_set_early_exit(eng.ctx, ptr(lane["bufs"]["n_done"]), n)
            try:
                for t in range(1, self.max_len):
                    for lane, stream in zip(lanes, streams):
                        with torch.cuda.stream(stream):
                            lane["eng"].decode_step(t, lane["n"], K, lane["enc"], lane["kv"], lane["bufs"], lane["bst"],
                                                    akv=lane["akv"])
            finally:
                for lane in lanes:
                    lane["eng"].lib.care_ctx_set_early_exit(lane["eng"].ctx, None, 0)
            for lane, stream in zip(lanes, streams):
                eng, outs = lane["eng"], lane["outs"]
                with torch.cuda.stream(stream):
                    check(eng.lib.care_beam_finalize(eng.ctx, ctypes.byref(lane["bst"]), beam_alpha, topk, ptr(outs[0]),
                                                     ptr(outs[1]), ptr(outs[2]), ptr(outs[3]), eng._stream()),
                          "care_beam_finalize")
            for side in streams[1:]:
                main.wait_stream(side)

        def launches():
            return sum(int(e.lib.care_ctx_launch_count(e.ctx)) for e in engines)

        key = (B, K, topk, beam_alpha, bos, n_lanes)
        entry = self._graphs.get(key)
        if entry is None:
            before = launches()
            body()          # eager pass: allocates every workspace, encodes the TMA descriptors, sets attributes
            n_launch = launches() - before
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                body()
            # the workspaces the captured launches point into: kept fresh on every replay (see _buf)
            used = [[k for k, e in eng._ws_epoch_of.items() if e == self._epoch] for eng in engines]
            self._graphs[key] = (graph, n_launch, used)
            self._graph_launches -= n_launch   # the capture pass went through the library's counter without running
            # the eager pass already produced this call's result
        else:
            for eng, keys in zip(engines, entry[2]):
                for k in keys:
                    eng._ws_epoch_of[k] = self._epoch
            entry[0].replay()
            self._graph_launches += entry[1]
        if n_lanes == 1:
            return tuple(o.clone() for o in lanes[0]["outs"])
        return tuple(torch.cat([lane["outs"][j] for lane in lanes], dim=0) for j in range(4))

    # ------------------------------------------------------------------------------------------
    # full-sequence decoder pass (mask-predict passes and the stateless decoding_phase)
    # ------------------------------------------------------------------------------------------
    def _attr_block_seq(self, x_in, x_out, akv, n_videos, rpv, N, x_in32=None, x_out32=None):
        lib, ctx, dt, w, d, T = self.lib, self.ctx, self.dt, self.w, self.d, self.tdtype
        st = self._stream()
        qa = self._buf("sq_qa", (N, d), T); cxa = self._buf("sq_ctx_a", (N, d), T)
        y32 = None if self._fused_for(N) else self._buf("sq_y32", (N, d), torch.float32)
        self.gemm(x_in, w["Waq"], w["baq"], qa, N, d, d)
        check(lib.care_group_attn(ctx, dt, ptr(qa), d, ptr(akv), 2 * d, 0, d, n_videos, rpv, akv.shape[1], self.H, d,
                                  None, 0, None, ptr(cxa), st), "care_group_attn(attr)")
        self._sublayer_tail(cxa, w["Wao"], w["bao"], w["lna_g"], w["lna_b"], x_in, x_out, N, d, y32, x_in32, x_out32)

    def _sequence_hidden(self, tokens, positions, R, L, kv, n_videos, causal, add_feats, gsg, akv=None):
        """Decoder layer over R sequences of L tokens (reference: Decoder/Transformer.py:161-237).
        tokens / positions: int32 [R*L]; kv: cross K/V [n_videos, Lm, 2d]; the R rows are video-major
        (R / n_videos consecutive rows per video).  Returns the hidden states [R*L, d]."""
        lib, ctx, dt, w, d, T = self.lib, self.ctx, self.dt, self.w, self.d, self.tdtype
        st = self._stream()
        N = R * L
        rpv = (R // n_videos) * L
        x0 = self._buf("sq_x0", (N, d), T); x1 = self._buf("sq_x1", (N, d), T)
        x2 = self._buf("sq_x2", (N, d), T); x3 = self._buf("sq_x3", (N, d), T)
        cx = self._buf("sq_ctx", (N, d), T); qc = self._buf("sq_qc", (N, d), T)
        qkv = self._buf("sq_qkv", (N, 3 * d), T)
        y32 = None if self._fused_for(N) else self._buf("sq_y32", (N, d), torch.float32)
        hb = self._buf("sq_ffn", (N, self.F), T)
        r0 = r1 = r2 = r3 = ra = None
        if self.fused_ln == 2:
            r0 = self._buf("sq_r0", (N, d), torch.float32); r1 = self._buf("sq_r1", (N, d), torch.float32)
            r2 = self._buf("sq_r2", (N, d), torch.float32); r3 = self._buf("sq_r3", (N, d), torch.float32)
        check(lib.care_embed_ln(ctx, dt, ptr(tokens), ptr(positions), 0, ptr(w["word"]), ptr(w["pos"]), ptr(add_feats),
                                ptr(gsg), rpv, ptr(w["emb_g"]), ptr(w["emb_b"]), self.eps, N, d, ptr(x0), ptr(r0), st),
              "care_embed_ln")
        self.gemm(x0, w["Wqkv"], w["bqkv"], qkv, N, 3 * d, d)
        check(lib.care_group_attn(ctx, dt, ptr(qkv), 3 * d, ptr(qkv), 3 * d, d, 2 * d, R, L, L, self.H, d, ptr(tokens),
                                  1 if causal else 0, None, ptr(cx), st), "care_group_attn(self)")
        self._sublayer_tail(cx, w["Wo"], w["bo"], w["ln1_g"], w["ln1_b"], x0, x1, N, d, y32, r0, r1)
        if self.attr_pos == "attr2cross":
            xa = self._buf("sq_xa", (N, d), T)
            ra = self._buf("sq_ra", (N, d), torch.float32) if self.fused_ln == 2 else None
            self._attr_block_seq(x1, xa, akv, n_videos, rpv, N, r1, ra)
            x1, r1 = xa, ra
        self.gemm(x1, w["Wxq"], w["bxq"], qc, N, d, d)
        check(lib.care_group_attn(ctx, dt, ptr(qc), d, ptr(kv), 2 * d, 0, d, n_videos, rpv, self.Lm, self.H, d, None, 0,
                                  ptr(w["hybrid_bias"]), ptr(cx), st), "care_group_attn(cross)")
        if self.attr_pos == "parallel":   # Layers.py:188-201
            qa = self._buf("sq_qa", (N, d), T); cxa = self._buf("sq_ctx_a", (N, d), T)
            cat = self._buf("sq_ctx_cat", (N, 2 * d), T)
            self.gemm(x1, w["Waq"], w["baq"], qa, N, d, d)
            check(lib.care_group_attn(ctx, dt, ptr(qa), d, ptr(akv), 2 * d, 0, d, n_videos, rpv, akv.shape[1], self.H, d,
                                      None, 0, None, ptr(cxa), st), "care_group_attn(attr)")
            torch.cat([cx, cxa], dim=1, out=cat)
            self._sublayer_tail(cat, w["Wpo"], w["bpo"], w["lnp_g"], w["lnp_b"], x1, x2, N, 2 * d, y32, r1, r2)
        else:
            self._sublayer_tail(cx, w["Wxo"], w["bxo"], w["ln2_g"], w["ln2_b"], x1, x2, N, d, y32, r1, r2)
        if self.attr_pos == "cross2attr":
            xa = self._buf("sq_xa", (N, d), T)
            ra = self._buf("sq_ra", (N, d), torch.float32) if self.fused_ln == 2 else None
            self._attr_block_seq(x2, xa, akv, n_videos, rpv, N, r2, ra)
            x2, r2 = xa, ra
        self.gemm(x2, w["W1"], w["b1"], hb, N, self.F, d, act=ACT_RELU)
        self._sublayer_tail(hb, w["W2"], w["b2"], w["ln3_g"], w["ln3_b"], x2, x3, N, self.F, y32, r2, r3)
        return x3

    def _memory_mean(self, memory):
        B = memory.shape[0]
        out = torch.empty((B, self.d), dtype=torch.float32, device=self.device)
        check(self.lib.care_rows_mean(self.ctx, self.dt, ptr(memory), B, memory.shape[1], self.d, ptr(out),
                                      self._stream()), "care_rows_mean")
        return out

    def sequence_logits(self, input_ids, inputs, last_only=False, decoding_type=None):
        """Stateless `decoding_phase` (Framework.py:240-269): logits for a whole batch of prefixes.
        `inputs['encoder_hidden_states']` may hold one row per video or be already repeated per beam
        (auto_enlarge): the cross K/V are projected for the rows that are given."""
        decoding_type = decoding_type or self.opt["decoding_type"]
        memory = inputs["encoder_hidden_states"]
        if memory.dtype != self.tdtype or not memory.is_contiguous():
            memory = memory.to(self.tdtype).contiguous()
        n_mem = memory.shape[0]
        R, L = input_ids.shape
        if R % n_mem != 0:
            raise ValueError("input_ids rows (%d) must be a multiple of the memory rows (%d)" % (R, n_mem))
        if L > self.max_len:
            raise ValueError("sequence length %d exceeds max_len %d" % (L, self.max_len))
        tokens = input_ids.to(self.device, torch.int32).contiguous().view(-1)
        positions = torch.arange(L, dtype=torch.int32, device=self.device).repeat(R)
        gsg = inputs.get("semantic_hidden_states") if self.use_gsg else None
        if gsg is not None:
            gsg = gsg.to(self.device, torch.float32).contiguous()
            if gsg.shape[0] != n_mem:
                raise ValueError("semantic_hidden_states rows must match encoder_hidden_states rows")
        nar = decoding_type == "NARFormer"
        add = self._memory_mean(memory) if nar else None
        kv = self.cross_kv(memory)
        akv = self.attr_kv(inputs)
        if akv is not None and akv.shape[0] != n_mem:
            raise ValueError("semantic_embs rows must match encoder_hidden_states rows")
        x3 = self._sequence_hidden(tokens, positions, R, L, kv, n_mem, not nar, add, gsg, akv)
        if last_only:
            x3 = x3.view(R, L, self.d)[:, -1, :].contiguous()
            rows = R
        else:
            rows = R * L
        logits = torch.empty((rows, self.ldv), dtype=torch.float32, device=self.device)
        self.gemm(x3, self.w["Wvocab"], None, logits, rows, self.V, self.d)
        logits = logits[:, :self.V]
        return logits if last_only else logits.view(R, L, self.V)

    # ------------------------------------------------------------------------------------------
    # mask-predict  (reference: models/Translator.py:240-318 + misc/Decoding/na_algorithms.py:146-197)
    # ------------------------------------------------------------------------------------------
    def _teacher_product(self, teacher, tokens, lengths, probs, R, L, n_len, out):
        """out[r,p] = probs[r,p] * p_teacher(y_p | y_<p) (scoring_by_teacher, na_algorithms.py:92-126): one
        teacher-forced pass of the auto-regressive teacher over [<bos>, y_0 .. y_{L-2}], in chunks of videos so that
        the fp32 logits of a chunk stay below ~2 GB."""
        t_eng, t_enc, mapping = teacher["engine"], teacher["enc"], teacher.get("mapping")
        tok = tokens.view(R, L).long()
        if mapping is not None:
            tok = mapping[tok]
        targets = tok.to(torch.int32).contiguous()
        ids = torch.cat([torch.full((R, 1), BOS, dtype=torch.int64, device=self.device), tok[:, :-1]], dim=1)
        B = R // n_len
        per_video = n_len * L * t_eng.ldv * 4
        step = max(1, min(B, (2 << 30) // per_video))
        for a in range(0, B, step):
            b = min(B, a + step)
            inputs = {k: v[a:b] for k, v in t_enc.items() if k in ("encoder_hidden_states", "semantic_hidden_states",
                                                                   "semantic_embs")}
            logits = t_eng.sequence_logits(ids[a * n_len:b * n_len], inputs, decoding_type="ARFormer")
            rows = slice(a * n_len, b * n_len)
            check(self.lib.care_nar_teacher_probs(self.ctx, ptr(logits), logits.stride(1), ptr(targets[rows]),
                                                  ptr(lengths.view(-1)[rows]), (b - a) * n_len, L, t_eng.V,
                                                  ptr(probs[rows]), ptr(out[rows]), self._stream()),
                  "care_nar_teacher_probs")
        return out

    def mask_predict(self, enc, opt, length_beam_size, length_bias=0, beam_alpha=1.0, trace=None, teacher=None):
        """`teacher` (optional): dict(engine, enc, mapping, masking, final) - the auto-regressive model whose
        probabilities multiply the student's when positions are re-masked (`masking`, opt masking_decision) and
        when the length candidates are ranked (`final`, opt no_candidate_decision = False)."""
        lib, ctx = self.lib, self.ctx
        st = self._stream()
        memory = enc["encoder_hidden_states"]
        B = memory.shape[0]
        i32 = torch.int32
        if "preds_length_logits" in enc:
            n = length_beam_size
            lengths = torch.empty((B, n), dtype=i32, device=self.device)
            lg = enc["preds_length_logits"]
            check(lib.care_nar_length_beam(ctx, ptr(lg), lg.stride(0), B, self.max_len, n, int(length_bias), 4,
                                           self.max_len, ptr(lengths), st), "care_nar_length_beam")
        else:
            lo, hi = opt.get("na_length_range", [5, 11])
            lengths = torch.arange(lo, hi, dtype=i32, device=self.device).unsqueeze(0).repeat(B, 1).contiguous()
            n = lengths.shape[1]
        L = int(lengths.max().item())   # the one host read before the passes (reference: Translator.py:273)
        R = B * n
        use_ct = bool(opt.get("use_ct", False))
        T = opt.get("iterations", 5) + (1 if use_ct else 0)
        # floor(len * (1 - c/T)) exactly as na_algorithms.py:176 computes it (fp32 tensor times python float)
        table = torch.stack([(torch.arange(self.max_len + 1).float() * (1.0 - (c / T))).long()
                             for c in range(T)]).to(self.device, i32).contiguous()
        tokens = torch.empty((R, L), dtype=i32, device=self.device)
        positions = torch.empty((R * L,), dtype=i32, device=self.device)
        probs = torch.empty((R, L), dtype=torch.float32, device=self.device)
        new_idx = torch.empty((R * L,), dtype=i32, device=self.device)
        new_prob = torch.empty((R * L,), dtype=torch.float32, device=self.device)
        mask_ind = torch.empty((R, L), dtype=torch.uint8, device=self.device)
        check(lib.care_nar_init(ctx, ptr(lengths), R, L, VIS if use_ct else MASK, ptr(tokens), ptr(positions),
                                ptr(probs), st), "care_nar_init")
        kv = self.cross_kv(memory)
        akv = self.attr_kv(enc)
        add = self._memory_mean(memory)
        gsg = enc.get("semantic_hidden_states") if self.use_gsg else None
        N = R * L

        def one_pass():
            x3 = self._sequence_hidden(tokens.view(-1), positions, R, L, kv, B, False, add, gsg, akv)
            if self.fused_vocab:
                nseg = int(lib.care_vocab_beam_nseg(ctx, N, self.V))
                part = self._buf("nar_partials", (N, nseg, 6), torch.float32)
                check(lib.care_vocab_beam_partials(ctx, ptr(x3), self.d, ptr(self.w["Wvocab"]),
                                                   self.w["Wvocab"].stride(0), N, self.V, self.d, 1, ptr(part), nseg,
                                                   st), "care_vocab_beam_partials")
                check(lib.care_nar_best_partials(ctx, ptr(part), nseg, N, self.V, ptr(new_idx), ptr(new_prob), st),
                      "care_nar_best_partials")
            else:
                logits = self._buf("nar_logits", (N, self.ldv), torch.float32)
                self.gemm(x3, self.w["Wvocab"], None, logits, N, self.V, self.d)
                check(lib.care_nar_best_logits(ctx, ptr(logits), self.ldv, N, self.V, ptr(new_idx), ptr(new_prob), st),
                      "care_nar_best_logits")

        one_pass()
        check(lib.care_nar_apply(ctx, ptr(tokens), ptr(probs), ptr(new_idx), ptr(new_prob), None, ptr(lengths), R, L,
                                 1 if use_ct else 0, st), "care_nar_apply")
        if trace is not None:
            trace.append(dict(c=0, tokens=tokens.cpu().clone(), probs=probs.cpu().clone()))
        scored = torch.empty((R, L), dtype=torch.float32, device=self.device) if teacher is not None else None
        for c in range(1, T):
            mode = 0 if (use_ct and c == 1) else 1
            worst = probs
            if teacher is not None and teacher.get("masking") and mode == 1:
                worst = self._teacher_product(teacher, tokens, lengths, probs, R, L, n, scored)
            check(lib.care_nar_remask(ctx, ptr(tokens), ptr(worst), ptr(lengths), ptr(table[c]), mode, R, L,
                                      ptr(mask_ind), st), "care_nar_remask")
            one_pass()
            check(lib.care_nar_apply(ctx, ptr(tokens), ptr(probs), ptr(new_idx), ptr(new_prob), ptr(mask_ind),
                                     ptr(lengths), R, L, 0, st), "care_nar_apply")
            if trace is not None:
                trace.append(dict(c=c, tokens=tokens.cpu().clone(), probs=probs.cpu().clone(),
                                  mask=mask_ind.cpu().clone()))
        out_tok = torch.empty((B, 1, L), dtype=i32, device=self.device)
        out_lp = torch.empty((B, 1, L), dtype=torch.float32, device=self.device)
        best = torch.empty((B,), dtype=i32, device=self.device)
        if teacher is not None and teacher.get("final", True):
            probs = self._teacher_product(teacher, tokens, lengths, probs, R, L, n, scored)
        check(lib.care_nar_select(ctx, ptr(tokens), ptr(probs), ptr(lengths), B, n, L, float(beam_alpha), ptr(out_tok),
                                  ptr(out_lp), ptr(best), st), "care_nar_select")
        if trace is not None:
            trace.append(dict(lengths=lengths.cpu().clone(), best=best.cpu().clone()))
        return out_tok, out_lp


def ensemble_ar_decode(engines, encs, B, beam_size=5, topk=1, beam_alpha=1.0, bos=None):
    """Beam search over the mean of the models' log-probabilities (reference: models/Translator.py:39-52,
    111-133).  Every model keeps its own KV cache and cross K/V; the beam state (tokens, ancestry, scores)
    is shared.  Logits are materialised per model (no fused vocabulary kernel on this path)."""
    e0 = engines[0]
    K = beam_size
    need = max(K, topk)
    lib = e0.lib
    st = e0._stream()
    for e in engines[1:]:
        if (e.V, e.max_len, e.device) != (e0.V, e0.max_len, e0.device):
            raise ValueError("ensembled models must share the vocabulary, max_len and device")
    bos = BOS if bos is None else int(bos)
    kvs = [e.cross_kv(enc["encoder_hidden_states"]) for e, enc in zip(engines, encs)]
    akvs = [e.attr_kv(enc) for e, enc in zip(engines, encs)]
    bufs, bst = e0._beam_buffers(B, K, need)
    check(lib.care_beam_init(e0.ctx, ctypes.byref(bst), bos, st), "care_beam_init")
    R = B * K
    logits = [e._buf("ens_logits", (R, e.ldv), torch.float32) for e in engines]
    mean_lp = e0._buf("ens_mean", (R, e0.ldv), torch.float32)
    ptrs = (ctypes.c_void_p * len(engines))(*[ptr(x) for x in logits])
    for t in range(1, e0.max_len):
        for e, enc, kv, akv, lg in zip(engines, encs, kvs, akvs, logits):
            x3 = e.step_hidden(t, B, K, enc, kv, bufs, akv)
            e.gemm(x3, e.w["Wvocab"], None, lg, R, e.V, e.d)
        check(lib.care_ensemble_logprobs(e0.ctx, ptrs, len(engines), e0.ldv, R, e0.V, ptr(mean_lp), st),
              "care_ensemble_logprobs")
        check(lib.care_beam_step_logprobs(e0.ctx, ctypes.byref(bst), ptr(mean_lp), e0.ldv, t, e0.max_len, None, None,
                                          st), "care_beam_step_logprobs")
    Tm = e0.max_len - 1
    out_tok = torch.empty((B, topk, Tm), dtype=torch.int32, device=e0.device)
    out_len = torch.empty((B, topk), dtype=torch.int32, device=e0.device)
    out_score = torch.empty((B, topk), dtype=torch.float32, device=e0.device)
    out_t = torch.empty((B, topk), dtype=torch.int32, device=e0.device)
    check(lib.care_beam_finalize(e0.ctx, ctypes.byref(bst), float(beam_alpha), topk, ptr(out_tok), ptr(out_len),
                                 ptr(out_score), ptr(out_t), st), "care_beam_finalize")
    return out_tok, out_len, out_score, out_t


def hyps_from_device(out_tok, out_len, out_score, out_t, beam_alpha, topk):
    """Host-side tail of Translator.collect_hypothesis_and_scores (models/Translator.py:211-220):
    scores are `float(score) / t ** alpha` in Python double precision (Beam.py:91-101), and the
    reference's n_best carry-over between videos (Translator.py:215) is reproduced."""
    tok = out_tok.cpu().tolist()
    ln_t = out_len.cpu()
    ln = ln_t.tolist()
    sc = out_score.cpu().tolist()
    tt = out_t.cpu().tolist()
    # n_best[v] = min(topk, min over u <= v of the number of finished hypotheses of video u)
    n_best = torch.cummin((ln_t > 0).sum(dim=1).clamp(max=topk), dim=0).values.tolist() if ln else []
    if topk == 1 and (not n_best or n_best[-1] == 1):
        hyps = [[t[0][:l[0]]] for t, l in zip(tok, ln)]
        scores = [[s[0] / t[0] ** beam_alpha] for s, t in zip(sc, tt)]
        return hyps, scores
    hyps, scores = [], []
    for v, nb in enumerate(n_best):
        hyps.append([tok[v][r][:ln[v][r]] for r in range(nb)])
        scores.append([sc[v][r] / tt[v][r] ** beam_alpha for r in range(nb)])
    return hyps, scores


def carry_n_best(hyps, scores, topk):
    """Re-applies the reference's running `n_best = min(n_best, finished hypotheses)` (Translator.py:215)
    across lists that were built per chunk: a video with fewer than `topk` hypotheses truncates every
    later video, also those of later chunks."""
    n_best = topk
    for v in range(len(hyps)):
        if len(hyps[v]) < n_best:
            n_best = len(hyps[v])
        elif len(hyps[v]) > n_best:
            hyps[v], scores[v] = hyps[v][:n_best], scores[v][:n_best]
    return hyps, scores
