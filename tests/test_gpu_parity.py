"""End-to-end parity of the CUDA path (through the reference-shaped API and the C ABI) against
(1) the golden outputs of the unmodified reference (tests/golden/*.json) and (2) the CPU oracle run
on the same seeded inputs.  fp32 mode: token sequences and concept ids exact (any mismatch must be
explained by an oracle decision margin below 1e-4, i.e. an fp32 summation-order tie).  16-bit mode
(fp16 operands, the default and benchmarked mode): logits within 1e-2 relative and per-step log-probs
within 1e-3 (scaled by the logit magnitude) of the oracle - the tolerances of BASELINE.json `north_star` -
and the measured sequence exact-match on 512 videos with trained-like weights (see
test_h16_exact_match_512 and DESIGN.md section 5 for what that number is and is not)."""
import os

import pytest
import torch

from oracle import care_oracle as co
from tests.helpers import load_golden, rebuild_case

pytestmark = pytest.mark.gpu

AR_CASES = ["cfg1_plain", "cfg1_sharp", "cfg2_plain", "cfg2_sharp", "cfg2_sharp_k3_nbest3_a07",
            "cfg2_sharp_greedy", "cfg3_sharp", "cfg4_sharp", "cab_sharp", "cab_attr2cross_sharp", "cab_parallel_sharp",
            "cfg4_plain", "cfg3_trained", "cfg4_trained"]


def _round_weights(sd, precision):
    """The oracle's weights for a 16-bit comparison: the GEMM weight matrices as the engine stores them."""
    if precision == "fp32":
        return sd
    dt = torch.float16 if precision.startswith("fp16") else torch.bfloat16
    return {k: (v.to(dt).float() if v.dim() == 2 and "embeddings" not in k and "hybrid_bias" not in k else v)
            for k, v in sd.items()}


def _gpu_model(opt, sd, precision):
    import care_b200
    m = care_b200.get_framework(dict(opt, care_precision=precision))
    m.load_state_dict(sd, strict=True)
    return m.eval().to("cuda")


def _oracle_margins(sd, opt, feats):
    hyps, scores, tr = co.ar_translate(sd, opt, feats, return_trace=True)
    margins = []
    for b in tr["beams"]:
        m = 1e9
        for rec in b.trace:
            vals = torch.cat([rec["scores"], torch.tensor([rec["runner_up"]])])
            m = min(m, float((vals[:-1] - vals[1:]).abs().min()))
        margins.append(m)
    return hyps, scores, margins, tr


@pytest.mark.parametrize("name", AR_CASES)
def test_fp32_matches_reference_golden(name):
    import care_b200
    rec = load_golden(name)
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    enc = model.encoding_phase([f.cuda() for f in feats])
    torch.cuda.synchronize()
    tie_videos = set()
    o_enc = co.encoding_phase(sd, opt, feats)
    mem_err = (enc["encoder_hidden_states"].cpu() - o_enc["encoder_hidden_states"]).abs().amax(dim=(1, 2))
    if "semantic_labels" in rec:
        p = o_enc["preds_attr"]
        srt = p.sort(dim=1, descending=True)[0]
        k = opt["use_attr_topk"]
        gaps = (srt[:, :k] - srt[:, 1:k + 1]).min(dim=1)[0]
        labels = enc["semantic_labels"].cpu()
        for v in range(labels.shape[0]):
            if labels[v].tolist() != rec["semantic_labels"][v]:
                # only an fp32 summation-order tie may reorder concepts
                assert gaps[v] < 1e-6, "concept ids differ with a clear margin %g (video %d)" % (gaps[v], v)
                tie_videos.add(v)
        assert (enc["preds_attr"].cpu() - p).abs().max().item() < 2e-6
        if "semantic_hidden_states" in o_enc:
            assert (enc["semantic_hidden_states"].cpu() - o_enc["semantic_hidden_states"]).abs().max().item() < 1e-4
        if "semantic_embs" in o_enc:
            sem_err = (enc["semantic_embs"].float().cpu() - o_enc["semantic_embs"]).abs().amax(dim=(1, 2))
            assert all(float(sem_err[i]) < 1e-4 for i in range(len(sem_err)) if i not in tie_videos)
    for v in range(mem_err.shape[0]):
        if v not in tie_videos:
            assert mem_err[v] < 1e-4, (v, float(mem_err[v]))
    hyps, scores = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    _, _, margins, _ = _oracle_margins(sd, opt, feats)
    exact = 0
    for v in range(len(hyps)):
        if hyps[v] == rec["hyps"][v]:
            exact += 1
            for a, b in zip(scores[v], rec["scores"][v]):
                assert v in tie_videos or abs(a - b) < 1e-4 * max(1.0, abs(b)), (v, a, b)
        elif v not in tie_videos:
            assert margins[v] < 1e-4, "video %d differs although the oracle margin is %g" % (v, margins[v])
    near_tie = [v for v in range(len(hyps)) if hyps[v] != rec["hyps"][v]]
    print("\n%s fp32: %d/%d sequences identical to the reference (concept-tie videos: %s, near-tie videos: %s)" % (
        name, exact, len(hyps), sorted(tie_videos), near_tie))
    # every video is exact unless it sits on an fp32 summation-order tie (audited above); at most one such video
    assert exact >= len(hyps) - max(1, len(tie_videos)), "only %d/%d sequences exact" % (exact, len(hyps))


def _prefixes_from_trace(step_rec, B, K):
    """Reconstructs the [B*K, t] input_ids the reference would have fed at this step."""
    t = step_rec["step"]
    anc, hist = step_rec["pre"]["anc"], step_rec["pre"]["tok_hist"]
    rows = []
    for v in range(B):
        for b in range(K):
            row = [int(hist[v, p, int(anc[v, b, p])]) for p in range(t - 1)] + [int(hist[v, t - 1, b])]
            rows.append(row)
    return torch.tensor(rows, dtype=torch.long)


@pytest.mark.parametrize("precision", ["fp32", "fp16", "fp16-stream", "bf16"])
@pytest.mark.parametrize("name", ["cfg2_sharp", "cfg1_sharp", "cfg4_sharp", "cab_sharp", "cab_parallel_sharp",
                                  "cfg3_trained", "cfg4_trained", "cfg4_plain"])
def test_teacher_forced_step_logits(name, precision):
    """Every step's logits from the KV-cached, ancestry-indirected CUDA path against the oracle's
    full-prefix recompute on exactly the prefixes the GPU beam holds at that step.  "bf16-stream" forces the
    live-slot chunk-stream self-attention kernel, which small batches would not select by themselves."""
    import care_b200
    rec = load_golden(name)
    opt, sd, feats = rebuild_case(rec, batch=3)
    sd_o = _round_weights(sd, precision)
    if precision == "fp16-stream":
        opt, precision = dict(opt, care_self_compact=3), "fp16"
    model = _gpu_model(opt, sd, precision)
    eng = model.engine()
    enc = model.encoding_phase([f.cuda() for f in feats])
    B, K = feats[0].shape[0], opt["beam_size"]
    trace = []
    eng.ar_decode(enc, B, beam_size=K, topk=1, trace=trace, trace_logits=True, early_exit_every=0)
    # decoder-only check: the oracle gets the GPU's own encode outputs (concept ranking is a separate,
    # tie-sensitive test), so every difference below comes from the per-step decoder kernels
    inputs = {k: co.repeat_rows(enc[k].float().cpu(), K) for k in co.decoder_input_keys(opt)}
    worst_logit, worst_lp = 0.0, 0.0
    for step_rec in trace:
        if step_rec["step"] not in (1, 2, 3, 5, 9, 17, 29):
            continue
        ids = _prefixes_from_trace(step_rec, B, K)
        ref = co.decoding_phase(sd_o, opt, ids, inputs, last_time_step_logits=True)
        got = step_rec["logits"]
        live = (step_rec["pre"]["done"] == 0).repeat_interleave(K)
        if step_rec["step"] == 1:
            live = live & (torch.arange(B * K) % K == 0)
        if not live.any():
            continue
        scale = ref[live].abs().max().item()
        worst_logit = max(worst_logit, (got[live] - ref[live]).abs().max().item() / scale)
        lp_err = (torch.log_softmax(got[live], 1) - torch.log_softmax(ref[live], 1)).abs()
        # log-probs that matter for the beam: the top of the distribution
        top = torch.log_softmax(ref[live], 1) > -12
        worst_lp = max(worst_lp, (lp_err * top).max().item() / max(1.0, scale))
    print("\n%s %s: max rel logit err %.3e, max log-prob err (scaled) %.3e" % (name, precision, worst_logit, worst_lp))
    if precision == "fp32":
        assert worst_logit < 2e-5 and worst_lp < 2e-5
    elif precision == "fp16":
        assert worst_logit < 1e-2       # north_star: 16-bit logits within 1e-2 relative (measured: < 1e-3)
        assert worst_lp < 1e-3          # north_star: per-step log-probs within 1e-3
    else:
        assert worst_logit < 1e-2       # the bf16 build: 8x coarser operands, logits still within the 1e-2 bound
        assert worst_lp < 1e-2


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("name", ["cfg2_plain", "cfg2_sharp", "cfg3_sharp", "cab_sharp", "cab_parallel_sharp", "cfg3_trained",
                                  "cfg4_trained"])
def test_h16_sequences(name, precision):
    """16-bit modes end to end on the small goldens: concept ids equal the fp32 reference's (the encoder runs as
    split products), sequences against the fp32 reference golden; a mismatching video must have a small oracle
    decision margin relative to the mode's logit noise."""
    import care_b200
    rec = load_golden(name)
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, precision)
    if "semantic_labels" in rec:
        labels = model.encoding_phase([f.cuda() for f in feats])["semantic_labels"].cpu().tolist()
        p = co.encoding_phase(sd, opt, feats)["preds_attr"].sort(dim=1, descending=True)[0]
        k = opt["use_attr_topk"]
        gaps = (p[:, :k] - p[:, 1:k + 1]).min(dim=1)[0]
        for v in range(len(labels)):
            assert labels[v] == rec["semantic_labels"][v] or gaps[v] < 2e-5, (v, float(gaps[v]))
    tr = care_b200.get_translator(opt)
    hyps, scores = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    _, _, margins, _ = _oracle_margins(sd, opt, feats)
    exact = sum(int(hyps[v] == rec["hyps"][v]) for v in range(len(hyps)))
    print("\n%s %s: %d/%d sequences identical to the fp32 reference; margins of the rest: %s" % (
        name, precision, exact, len(hyps), ["%.2e" % margins[v] for v in range(len(hyps)) if hyps[v] != rec["hyps"][v]]))
    scale = 20.0 if "trained" in name else 6.0 if "sharp" in name else 1.0   # logit magnitude of the weight set
    noise = (1e-3 if precision == "fp16" else 1e-2) * scale                  # the mode's log-prob error bound
    for v in range(len(hyps)):
        if hyps[v] != rec["hyps"][v]:
            assert margins[v] < noise, "video %d differs although the oracle margin is %g" % (v, margins[v])
        else:
            assert abs(scores[v][0] - rec["scores"][v][0]) < 0.05 * max(1.0, abs(rec["scores"][v][0]))


@pytest.mark.parametrize("population", ["cfg4_trained_512", "cfg4_trained_512_greedy"])
@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_h16_exact_match_512(precision, population):
    """The north-star population: VATEX-large shape (cfg4), beam 5, 512 videos, trained-like peaked weights
    (oracle/weights.py TRAINED), 16-bit mode against the hypotheses of the UNMODIFIED reference run on the CPU in
    fp32 (tests/golden/cfg4_trained_512.json).  Writes the exact-match rate per oracle-margin bucket to
    gpurun_out/ (committed under profiles/).  The fp32 mode must reproduce the golden exactly (near-ties aside);
    the 16-bit rate is asserted at its measured level - see DESIGN.md section 5 for the error budget behind it."""
    import json
    import care_b200
    rec = load_golden(population)
    greedy = population.endswith("greedy")
    opt, sd, feats = rebuild_case(rec)
    dev = [f.cuda() for f in feats]
    tr = care_b200.get_translator(opt)
    margins = rec["oracle_min_margin"]
    out = {}
    for prec in ("fp32", precision):
        model = _gpu_model(opt, sd, prec)
        labels = model.encoding_phase(dev)["semantic_labels"].cpu().tolist()
        hyps, scores = tr.translate_batch([model], {"feats": dev})
        same = [hyps[v] == rec["hyps"][v] for v in range(len(hyps))]
        concept_same = sum(int(labels[v] == rec["semantic_labels"][v]) for v in range(len(hyps)))
        buckets = [0.0, 1e-4, 1e-3, 1e-2, 3e-2, 1e-1, 3e-1, 1e9]
        curve = []
        for lo, hi in zip(buckets[:-1], buckets[1:]):
            idx = [v for v in range(len(hyps)) if lo <= margins[v] < hi]
            curve.append(dict(margin_lo=lo, margin_hi=hi, videos=len(idx), exact=sum(int(same[v]) for v in idx)))
        out[prec] = dict(exact=sum(same), videos=len(hyps), concept_ids_exact=concept_same, by_oracle_margin=curve)
        print("\ncfg4 trained-like, 512 videos, %s: %d/%d sequences identical to the reference (%.2f%%), concept ids "
              "identical for %d videos" % (prec, sum(same), len(hyps), 100.0 * sum(same) / len(hyps), concept_same))
        for c in curve:
            print("   oracle margin [%.0e, %.0e): %d/%d" % (c["margin_lo"], c["margin_hi"], c["exact"], c["videos"]))
        if prec == "fp32":
            bad = [v for v in range(len(hyps)) if not same[v] and margins[v] >= 1e-4]
            assert not bad, "fp32 mode differs from the reference on videos %s with clear margins" % bad[:8]
            assert concept_same >= len(hyps) - 2
        else:
            lens = [len(h[0]) for h in rec["hyps"]]
            assert min(lens) <= 5 and max(lens) >= 25     # the population really spreads over short and long captions
            assert concept_same >= len(hyps) - 3          # split-product encoder: concept ranking as in fp32
            # measured (round 2, B200): beam 5 fp16 503/512 (98.2 %), bf16 440/512; greedy fp16 see profiles/
            floor = (0.97 if greedy else 0.95) if prec == "fp16" else 0.70
            assert sum(same) >= floor * len(hyps), "%s exact-match %d/%d" % (prec, sum(same), len(hyps))
            # a mismatch may only happen where the reference's own decision margin is within the mode's noise
            noise = (1e-3 if prec == "fp16" else 1e-2) * 20.0 * 3
            assert all(same[v] or margins[v] < noise for v in range(len(hyps)))
        del model
    os.makedirs("gpurun_out", exist_ok=True)
    with open(os.path.join("gpurun_out", "h16_exact_match_%s%s.json" % (precision, "_greedy" if greedy else "")), "w") as f:
        json.dump(out, f, indent=1)


@pytest.mark.parametrize("R,split", [(20480, 0), (20480, 1), (2560, 0), (2560, 1), (133, 1)])
def test_fused_vocab_records_direct(R, split):
    """The bench's largest single kernel checked directly (not through the unfused CUDA path): the
    (max, sum-exp, top-(K+1)) records of care_vocab_beam_partials at R = 20480 (one GPU) / 2560 (an 8-GPU shard),
    V = 14745 against fp32 torch logsumexp / topk of the same 16-bit inputs; both epilogue schedules (alternate tiles,
    column halves)."""
    import ctypes
    from care_b200 import _lib
    lib = _lib.load("fp16")
    h = ctypes.c_void_p()
    _lib.check(lib.care_ctx_create(ctypes.byref(h), 0), "ctx")
    try:
        V, d, K = 14745, 1024, 5
        _lib.check(lib.care_ctx_set_option(h, b"vocab_split", split), "option")
        g = torch.Generator(device="cuda").manual_seed(3)
        x = torch.randn(R, d, device="cuda", generator=g).half()
        W = (torch.randn(V, d, device="cuda", generator=g) * 0.08).half()
        nseg = int(lib.care_vocab_beam_nseg(h, R, V))
        kb = 6
        # nseg is the maximum over row blocks; segments a row block does not have stay untouched (beam.cu derives
        # which exist from the tile schedule), so they start as neutral records here
        part = torch.full((R, nseg, 2 + 2 * kb), float("-inf"), device="cuda")
        part[:, :, 1] = 0.0
        st = torch.cuda.current_stream().cuda_stream
        _lib.check(lib.care_vocab_beam_partials(h, x.data_ptr(), d, W.data_ptr(), d, R, V, d, K, part.data_ptr(), nseg,
                                                st), "vocab_beam")
        torch.cuda.synchronize()
        worst_lse = worst_top = 0.0
        for r0 in range(0, R, 2048):
            logits = x[r0:r0 + 2048].float() @ W.float().t()
            lse_ref = torch.logsumexp(logits, 1)
            top_ref, idx_ref = logits.topk(K + 1, dim=1)
            p = part[r0:r0 + 2048]
            m, se = p[:, :, 0], p[:, :, 1]
            gm = m.max(dim=1, keepdim=True)[0]
            lse = gm.squeeze(1) + torch.log((se * torch.exp(m - gm)).sum(1))
            vals = p[:, :, 2:2 + kb].reshape(p.shape[0], -1)
            ids = p[:, :, 2 + kb:2 + 2 * kb].reshape(p.shape[0], -1).contiguous().view(torch.int32)
            top, pos = vals.topk(K + 1, dim=1)
            worst_lse = max(worst_lse, (lse - lse_ref).abs().max().item())
            worst_top = max(worst_top, (top - top_ref).abs().max().item())
            got_idx = ids.gather(1, pos)
            # ids must match wherever the reference's adjacent candidates are separated by more than the error
            gap = (top_ref[:, :-1] - top_ref[:, 1:]).min(dim=1)[0]
            clear = gap > 1e-3
            assert torch.equal(got_idx[clear].long(), idx_ref[clear]), "candidate ids differ on rows with clear gaps"
        print("\nfused vocab records at R=%d V=%d: max |lse err| %.2e, max |top-(K+1) value err| %.2e" % (
            R, V, worst_lse, worst_top))
        assert worst_lse < 2e-4 and worst_top < 2e-4
    finally:
        lib.care_ctx_destroy(h)


def test_wrapper_checkpoint_roundtrip(tmp_path):
    """Lightning-layout checkpoint -> load_model (reference defaults: data paths rewritten) -> translate_step, as
    translate.py drives it: the latency branch, then the non-latency branch whose teacher-forced pass feeds the
    concept criterion (models/Wrapper.py:182-184) - its mAP / F1 table must equal the criterion applied to the
    oracle's concept probabilities."""
    import pickle
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec, batch=4)
    old_base, new_base = os.path.join(str(tmp_path), "author"), os.path.join(str(tmp_path), "here")
    os.makedirs(os.path.join(new_base, "MSRVTT"))
    vocab = {i: "w%d" % i for i in range(opt["vocab_size"])}
    with open(os.path.join(new_base, "MSRVTT", "info_corpus.pkl"), "wb") as f:
        pickle.dump({"info": {"itow": vocab}}, f)
    opt = dict(opt, care_precision="fp32", dataset="MSRVTT",
               info_corpus=os.path.join(old_base, "MSRVTT", "info_corpus.pkl"),
               reference=os.path.join(old_base, "MSRVTT", "refs.pkl"),
               feats_a=[os.path.join(old_base, "MSRVTT", "feats", "audio.hdf5")])
    m = care_b200.Model(opt)
    m.captioner.load_state_dict(sd)
    path = os.path.join(str(tmp_path), "best.ckpt")
    torch.save(m.to_checkpoint(), path)
    model = care_b200.load_model(path, device=torch.device("cuda"), strict=True, base_data_path=new_base)
    assert model.get_opt()["feats_a"] == [os.path.join(new_base, "MSRVTT", "feats", "audio.hdf5")]
    assert model.get_vocab() == vocab          # read through the rewritten info_corpus path
    assert model.get_keys_to_device() == ["feats", "input_ids"]
    g = torch.Generator().manual_seed(5)
    labels = (torch.rand(4, 500, generator=g) < 0.03).float()
    labels[:, 3] = 1.0
    batch = {"feats": [f.cuda() for f in feats], "video_ids": ["video%d" % i for i in range(4)],
             "input_ids": torch.randint(4, opt["vocab_size"], (4, 9), generator=g).cuda(), "labels_attr": labels}
    out = model.translate_step(batch, vocab, assert_only_a_caption_per_video=True, inference_latency=True)
    assert model.eval_criterion.get_loss_info()["F1-05"] == 0      # latency branch: criterion untouched
    for i in range(4):
        item = out["video%d" % i][0]
        ref_caption = care_b200.to_sentence(rec["hyps"][i][0], vocab)
        assert item["caption"] == ref_caption and item["image_id"] == "video%d" % i
        assert abs(item["score"] - rec["scores"][i][0]) < 1e-4 * max(1.0, abs(rec["scores"][i][0]))
    out2 = model.translate_step(batch, vocab)                      # non-latency branch
    assert {k: v[0]["caption"] for k, v in out2.items()} == {k: v[0]["caption"] for k, v in out.items()}
    info = model.eval_criterion.get_loss_info()
    want = care_b200.get_criterion(opt, skip_crit_list=["lang"], override_opt={"calculate_mAP": True})
    want.get_loss({"preds_attr": co.encoding_phase(sd, opt, feats)["preds_attr"], "labels_attr": labels})
    for k, v in want.get_loss_info().items():
        assert abs(info[k] - v) < 1e-5 * max(1.0, abs(v)), (k, info[k], v)
    scores, _, preds = model.test_epoch_end([out2], verbose=False, keys_added_to_scores=["beam_size"], analyze=False)
    assert abs(scores["mAP"] - want.get_loss_info()["mAP"]) < 1e-5 and scores["beam_size"] == opt["beam_size"]
    assert set(preds) == {"video%d" % i for i in range(4)}
    assert model.eval_criterion.get_loss_info()["F1-05"] == 0      # recorders reset by the evaluation


def test_model_ensemble_wrapper(tmp_path):
    """ModelEnsemble (models/Wrapper.py:617-714) over two checkpoints: the wrapper's translate_step equals the
    reference's ensemble golden (mean of the models' log-probabilities)."""
    import care_b200
    rec = load_golden("ens2_cfg2_sharp")
    from tests.helpers import rebuild_ensemble_case
    opt, sds, feats = rebuild_ensemble_case(rec)
    opt = dict(opt, **{"feats_%s" % c: ["/data/feats_%s.hdf5" % c] for c in opt["modality"]})
    paths = []
    for i, sd in enumerate(sds):
        m = care_b200.Model(dict(opt, care_precision="fp32"))
        m.captioner.load_state_dict(sd)
        paths.append(os.path.join(str(tmp_path), "m%d.ckpt" % i))
        torch.save(m.to_checkpoint(), paths[-1])
    model = care_b200.load_model(paths, device=torch.device("cuda"), strict=True, replace_paths=False)
    assert isinstance(model, care_b200.ModelEnsemble) and isinstance(model.captioner, list) and len(model.captioner) == 2
    assert not model.need_to_split_feats
    vocab = {i: "w%d" % i for i in range(opt["vocab_size"])}
    B = feats[0].shape[0]
    batch = {"feats": [f.cuda() for f in feats], "video_ids": ["v%d" % i for i in range(B)]}
    out = model.translate_step(batch, vocab)
    for i in range(B):
        assert out["v%d" % i][0]["caption"] == care_b200.to_sentence(rec["hyps"][i][0], vocab)
        assert abs(out["v%d" % i][0]["score"] - rec["scores"][i][0]) < 1e-4 * max(1.0, abs(rec["scores"][i][0]))


def test_first_step_on_one_row_per_video_is_exact_and_ignores_stale_cache_memory():
    """Step 1 runs on ONE row per video in the 16-bit fused path (all K beams hold <bos>; engine.step_hidden
    compact_first): identical captions to the full-row first step, also when the KV-cache workspace is carved out of
    memory full of NaN bit patterns (only slot 0 of position 0 is written; the dense self-attention tile loads every
    slot of a position and 0 x NaN would poison the context)."""
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec)
    dev = [f.cuda() for f in feats]
    tr = care_b200.get_translator(opt)
    full = _gpu_model(dict(opt, care_compact_first_step=False), sd, "fp16")
    h_full, s_full = tr.translate_batch([full], {"feats": dev})
    assert not full.engine().compact_first
    del full
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    junk = [torch.full((32 << 20,), float("nan"), device="cuda", dtype=torch.float16) for _ in range(4)]
    torch.cuda.synchronize()
    del junk                                   # back to the caching allocator: the next engine's workspaces reuse it
    compact = _gpu_model(opt, sd, "fp16")
    h_c, s_c = tr.translate_batch([compact], {"feats": dev})
    assert compact.engine().compact_first
    assert h_c == h_full
    for a, b in zip(s_c, s_full):
        assert abs(a[0] - b[0]) < 2e-3 * max(1.0, abs(b[0]))     # one-row and K-row first steps take different GEMM tiles
    h_again, _ = tr.translate_batch([compact], {"feats": dev})    # graph replay
    assert h_again == h_c


@pytest.mark.parametrize("name,stream", [("cfg2_sharp", False), ("cfg4_trained", True), ("cab_sharp", False)])
def test_next_step_prologue_in_the_beam_kernel_changes_nothing(name, stream):
    """The beam kernel of step t writes step t+1's decoder input rows (care_ctx_set_next_step) and, for the chunk-stream
    self-attention, its live-slot records: same arithmetic as care_embed_ln / the record kernel, so captions AND scores
    are bit-identical to the run with the stand-alone launches - which it saves (one or two launches per step).  Both
    switches are off by default: measured neutral to slower."""
    import care_b200
    rec = load_golden(name)
    opt, sd, feats = rebuild_case(rec)
    if stream:
        opt = dict(opt, care_self_compact=3)      # the live-slot stream kernel for every shape
    dev = [f.cuda() for f in feats]
    tr = care_b200.get_translator(opt)
    out, launches = {}, {}
    for fuse in (False, True, "embed"):
        # "embed": the records ride on the step's own embedding launch (care_ctx_request_records; the default)
        fopt = dict(care_fuse_next_step=False, care_fuse_info=2) if fuse == "embed" else \
            dict(care_fuse_next_step=fuse, care_fuse_info=int(fuse))
        m = _gpu_model(dict(opt, care_cuda_graph=False, **fopt), sd, "fp16")
        n0 = m.engine().launch_count()
        out[fuse] = tr.translate_batch([m], {"feats": dev})
        launches[fuse] = m.engine().launch_count() - n0
    for fuse in (True, "embed"):
        assert out[fuse][0] == out[False][0]
        assert out[fuse][1] == out[False][1]
    assert launches[True] < launches[False]
    if stream:
        assert launches["embed"] < launches[False]
    print("\n%s: launches per decode %d -> %d (beam kernel) / %d (embedding launch)" % (
        name, launches[False], launches[True], launches["embed"]))


def test_no_gpu_no_fallback_message():
    from care_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH)


NAR_CASES = ["cfg5_plain", "cfg5_sharp", "cfg5_noct_sharp"]


@pytest.mark.parametrize("name", NAR_CASES)
def test_nar_fp32_matches_reference_golden(name):
    """Mask-predict (config 5) in the fp32 parity mode: chosen lengths, tokens and per-token log-probs
    against the unmodified reference's outputs; a differing video must be explained by an oracle
    near-tie (length-candidate score gap or token probability gap)."""
    import care_b200
    rec = load_golden(name)
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    hyps, lprobs = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    o_h, o_p, otr = co.nar_translate(sd, opt, feats, return_trace=True)
    assert o_h == rec["hyps"]
    enc = model.encoding_phase([f.cuda() for f in feats])
    if "preds_length" in rec:
        got = enc["preds_length"].cpu()
        assert (got - torch.tensor(rec["preds_length"])).abs().max().item() < 1e-4
    exact = 0
    for v in range(len(hyps)):
        if hyps[v] == rec["hyps"][v]:
            exact += 1
            a, b = torch.tensor(lprobs[v][0]), torch.tensor(rec["scores"][v][0])
            assert (a - b).abs().max().item() < 2e-4 * max(1.0, b.abs().max().item()), (v, a, b)
        else:
            top2 = otr["avg"][v].topk(2)[0]
            assert float(top2[0] - top2[1]) < 1e-3, "video %d differs with a clear candidate margin" % v
    print("\n%s fp32: %d/%d mask-predict outputs identical to the reference" % (name, exact, len(hyps)))
    assert exact >= 0.75 * len(hyps)


@pytest.mark.parametrize("precision", ["fp32", "fp16"])
@pytest.mark.parametrize("name", ["nar_cfg5_teacher_sharp", "nar_cfg5_teacher_masking_sharp", "nar_cfg5_teacher_mapped_sharp"])
def test_nar_teacher_rescoring(name, precision, tmp_path):
    """Mask-predict with an auto-regressive teacher rescoring the candidates (models/Translator.py:250-264,
    na_algorithms.py:92-126): final rescoring, per-iteration masking decisions, a teacher with another vocabulary
    order - against the reference's own outputs (fp32 mode: tokens exact unless the oracle's candidate margin is a
    near-tie; fp16: well-formed and mostly the same length)."""
    import pickle
    from types import SimpleNamespace
    import care_b200
    from tests.helpers import rebuild_teacher_case
    rec = load_golden(name)
    opt, sd, feats, teacher = rebuild_teacher_case(rec)
    words = {i: "w%d" % i for i in range(opt["vocab_size"])}
    mapping = teacher["vocab_mapping"]
    t_words = words if mapping is None else {int(mapping[i]): w for i, w in words.items()}
    t_opt = dict(teacher["opt"])
    for fn, vocab, o in (("student.pkl", words, opt), ("teacher.pkl", t_words, t_opt)):
        with open(os.path.join(str(tmp_path), fn), "wb") as f:
            pickle.dump({"info": {"itow": vocab}}, f)
        o["info_corpus"] = os.path.join(str(tmp_path), fn)
    model = _gpu_model(opt, sd, precision)
    t_model = _gpu_model(t_opt, teacher["sd"], precision)
    wrapper = SimpleNamespace(captioner=t_model, get_opt=lambda: t_opt)
    tr = care_b200.get_translator(opt)
    hyps, lprobs = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]}, teacher_model_wrapper=wrapper, vocab=words)
    assert (tr.vocab_mapping is None) == (mapping is None)
    _, _, otr = co.nar_translate(sd, opt, feats, return_trace=True, teacher=teacher)
    exact = 0
    for v in range(len(hyps)):
        assert len(hyps[v]) == 1 and len(hyps[v][0]) == len(lprobs[v][0])
        if hyps[v] == rec["hyps"][v]:
            exact += 1
            if precision == "fp32":
                a, b = torch.tensor(lprobs[v][0]), torch.tensor(rec["scores"][v][0])
                assert (a - b).abs().max().item() < 5e-4 * max(1.0, b.abs().max().item()), (v, a, b)
        elif precision == "fp32":
            top2 = otr["avg"][v].topk(2)[0]
            assert float(top2[0] - top2[1]) < 1e-3, "video %d differs with a clear candidate margin" % v
    print("\n%s %s: %d/%d teacher-rescored outputs identical to the reference" % (name, precision, exact, len(hyps)))
    if precision == "fp32":
        assert exact >= len(hyps) - 1
    # without the teacher the same model gives other log-probabilities: the rescoring really ran
    plain_h, plain_p = care_b200.get_translator(opt).translate_batch([model], {"feats": [f.cuda() for f in feats]})
    assert plain_p != lprobs


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
@pytest.mark.parametrize("name", ["cfg5_sharp", "cfg5_plain"])
def test_nar_h16(name, precision):
    """16-bit mask-predict: well-formed output, same chosen length for most videos, token agreement."""
    import care_b200
    rec = load_golden(name)
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, precision)
    tr = care_b200.get_translator(opt)
    hyps, lprobs = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    assert len(hyps) == len(rec["hyps"])
    agree, total = 0, 0
    for v in range(len(hyps)):
        assert len(hyps[v]) == 1 and len(hyps[v][0]) == len(lprobs[v][0])
        ref = rec["hyps"][v][0]
        got = hyps[v][0]
        n_ref = sum(1 for t in ref if t != 0)
        n_got = sum(1 for t in got if t != 0)
        if n_ref == n_got:
            total += n_ref
            agree += sum(int(a == b) for a, b in zip(got[:n_ref], ref[:n_ref]))
    print("\n%s %s: %d/%d tokens identical on same-length outputs" % (name, precision, agree, total))
    assert total > 0 and agree >= (0.9 if precision == "fp16" else 0.6) * total


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
@pytest.mark.parametrize("name", ["cfg2_sharp", "cfg5_sharp", "cab_attr2cross_sharp"])
def test_decoding_phase_stateless(name, precision):
    """Framework.decoding_phase(input_ids, inputs) (full prefix, no cache) against the oracle, for
    per-video and per-beam-row (auto_enlarge'd) memory layouts, all positions and last position."""
    rec = load_golden(name)
    opt, sd, feats = rebuild_case(rec, batch=3)
    model = _gpu_model(opt, sd, precision)
    enc = model.encoding_phase([f.cuda() for f in feats])
    inputs = model.prepare_inputs_for_decoder(enc, {})
    g = torch.Generator().manual_seed(5)
    rep, L = 2, 7
    ids = torch.randint(6, opt["vocab_size"], (3 * rep, L), generator=g)
    if opt["decoding_type"] == "ARFormer":
        ids[:, 0] = 2
    ids[1, 4] = 0
    ids[4, L - 1] = 0
    o_inputs = {k: co.repeat_rows(v.float().cpu(), rep) for k, v in inputs.items()}
    sd_o = _round_weights(sd, precision)
    ref_all = co.decoding_phase(sd_o, opt, ids, o_inputs)
    ref_last = co.decoding_phase(sd_o, opt, ids, o_inputs, last_time_step_logits=True)
    got_all = model.decoding_phase(ids.cuda(), inputs)["logits"].cpu()
    enlarged = {k: co.repeat_rows(v, rep) for k, v in inputs.items()}
    got_last = model.decoding_phase(ids.cuda(), enlarged, last_time_step_logits=True)["logits"].cpu()
    tol = {"fp32": 2e-5, "fp16": 2e-3, "bf16": 1e-2}[precision]
    scale = ref_all.abs().max().item()
    assert got_all.shape == ref_all.shape and got_last.shape == ref_last.shape
    assert (got_all - ref_all).abs().max().item() < tol * scale
    assert (got_last - ref_last).abs().max().item() < tol * scale


def test_host_batches_chunked_and_streamed_match_single_call():
    """Host-resident features: the chunk-pipelined translate_batch and the translate_stream generator
    return exactly what one device-resident call returns (videos are independent units)."""
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    ref_h, ref_s = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    tr.pipeline_chunk = 5
    pinned = [f.pin_memory() for f in feats]
    h, s = tr.translate_batch([model], {"feats": pinned})
    assert h == ref_h and s == ref_s
    batches = [{"feats": [f[a:b].contiguous().pin_memory() for f in feats]} for a, b in ((0, 4), (4, 8), (8, 12))]
    got = list(tr.translate_stream([model], batches))
    assert len(got) == 3
    assert sum((g[0] for g in got), []) == ref_h
    assert sum((g[1] for g in got), []) == ref_s
    assert list(tr.translate_stream([model], [])) == []


def test_chunked_host_batch_nbest_and_pageable_memory():
    """topk = 3 with a chunk size that splits the batch unevenly, from pageable (not pinned) host memory:
    the chunked call equals the golden of the unmodified reference."""
    import care_b200
    rec = load_golden("cfg2_sharp_k3_nbest3_a07")
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    tr.pipeline_chunk = 7
    h, s = tr.translate_batch([model], {"feats": [f.clone() for f in feats]})
    assert h == rec["hyps"]
    ref_h, ref_s = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    assert h == ref_h and s == ref_s


def test_workspace_limit_evicts_stale_sets_and_keeps_results():
    """A tiny workspace budget: every new batch size drops the workspaces (and CUDA graphs) of the previous
    sizes; results stay those of the reference and only the current call's buffers remain."""
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(dict(opt, care_workspace_limit_bytes=1), sd, "fp32")
    tr = care_b200.get_translator(opt)
    dev = [f.cuda() for f in feats]
    eng = model.engine()
    for n in (12, 5, 12, 3, 3, 12, 5):
        h, _ = tr.translate_batch([model], {"feats": [f[:n].contiguous() for f in dev]})
        assert h == rec["hyps"][:n]
        assert all(e == eng._epoch for e in eng._ws_epoch_of.values())
        assert set(eng._ws) == set(eng._ws_epoch_of)
        assert eng._ws_bytes == sum(t.numel() * t.element_size() for t in eng._ws.values())
    roomy = _gpu_model(opt, sd, "fp32")
    for n in (12, 5):
        tr.translate_batch([roomy], {"feats": [f[:n].contiguous() for f in dev]})
    assert len({k[1] for k in roomy.engine()._ws if k[0] == "bs_scores"}) == 2   # both sizes stay resident


@pytest.mark.parametrize("precision", ["fp32", "fp16", "bf16"])
def test_cuda_graph_replay_matches_eager(precision):
    """Small batches replay the whole decode as one CUDA graph: first call (eager + capture), replays,
    and a graph-free engine must all return the same hypotheses; replays count their launches."""
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec)
    dev = [f.cuda() for f in feats]
    tr = care_b200.get_translator(opt)
    eager = _gpu_model(dict(opt, care_cuda_graph=False), sd, precision)
    ref = tr.translate_batch([eager], {"feats": dev})
    model = _gpu_model(opt, sd, precision)
    first = tr.translate_batch([model], {"feats": dev})
    n0 = model.engine().launch_count()
    second = tr.translate_batch([model], {"feats": dev})
    n1 = model.engine().launch_count()
    other = [f[[3, 1, 2, 0, 5, 4, 7, 6, 9, 8, 11, 10]].contiguous() for f in dev]   # same shape, other inputs
    third = tr.translate_batch([model], {"feats": other})
    assert first == ref and second == ref
    assert n1 - n0 > 29 * 10, "graph replays must be counted as launches"
    perm = [3, 1, 2, 0, 5, 4, 7, 6, 9, 8, 11, 10]
    assert third[0] == [ref[0][i] for i in perm]
    if precision == "fp32":
        assert first[0] == rec["hyps"]


@pytest.mark.parametrize("over", [
    dict(beam_size=8, topk=4, max_len=14, vocab_size=1203),
    dict(beam_size=7, topk=1, max_len=30, vocab_size=9468, beam_alpha=0.5),
    dict(beam_size=2, topk=2, max_len=6, vocab_size=517),
])
def test_unusual_shapes_fp32_vs_oracle(over):
    """Beam widths up to the kernels' limit (8), n_best > 1, short max_len, odd vocabulary sizes: the fp32
    CUDA path against the oracle (no golden for these; the oracle itself is pinned to the reference)."""
    import care_b200
    from synth.shapes import CONFIGS, make_feats, make_opt
    from synth.weights import SHARP, make_state_dict
    opt = make_opt(**{**CONFIGS["cfg2"], **over})
    sd = make_state_dict(opt, seed=31, perturb=True, sharpen=SHARP)
    feats = make_feats(opt, 7, seed=13)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    hyps, scores = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    o_h, o_s, margins, _ = _oracle_margins(sd, opt, feats)
    exact = 0
    for v in range(len(hyps)):
        if hyps[v] == o_h[v]:
            exact += 1
            for a, b in zip(scores[v], o_s[v]):
                assert abs(a - b) < 1e-4 * max(1.0, abs(b)), (v, a, b)
        else:
            assert margins[v] < 1e-4, "video %d differs although the oracle margin is %g" % (v, margins[v])
    assert exact >= 5, "only %d/7 sequences identical" % exact


def test_single_video_and_many_videos_agree():
    """A video decodes to the same caption alone (batch 1, graph path) and inside a larger batch."""
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    for v in (0, 5, 11):
        h, s = tr.translate_batch([model], {"feats": [f[v:v + 1].cuda() for f in feats]})
        assert h[0] == rec["hyps"][v]


def test_empty_batch():
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    assert tr.translate_batch([model], {"feats": [f[:0].cuda() for f in feats]}) == ([], [])


def test_randomised_differential_fp32():
    """tests/fuzz_parity.py: random beam widths / n_best / max_len / vocab sizes / batch sizes / model
    families against the oracle; every mismatch must sit on an oracle near-tie."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "tests", "fuzz_parity.py"), "24", "3"],
                         capture_output=True, text=True, timeout=600)
    print(out.stdout[-400:])
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]


def test_full_size_batch_properties():
    """BASELINE.json's full size (cfg4, 4096 videos = 20480 beam rows per GPU), where no CPU oracle run is
    affordable for every video: size-independent properties.  (1) fp16: the second half of the batch is a
    copy of the first, so both halves must decode to identical captions (rows are independent; this
    exercises every kernel's addressing at the high row indices, the multi-segment vocabulary records and
    the CTA-pair GEMM tiles).  (2) fp32: eight videos spread over the batch match the CPU oracle and the
    same videos decoded alone."""
    import care_b200
    from synth.shapes import CONFIGS, make_feats, make_opt
    from synth.weights import SHARP, make_state_dict
    opt = make_opt(**CONFIGS["cfg4"])
    sd = make_state_dict(opt, seed=5, perturb=True, sharpen=SHARP)
    half = make_feats(opt, 2048, seed=77)
    feats = [torch.cat([f, f]).cuda() for f in half]
    tr = care_b200.get_translator(opt)
    m16 = _gpu_model(opt, sd, "fp16")
    hyps, scores = tr.translate_batch([m16], {"feats": feats})
    assert len(hyps) == 4096 and all(len(h) == 1 and 1 <= len(h[0]) <= 29 for h in hyps)
    assert hyps[:2048] == hyps[2048:] and scores[:2048] == scores[2048:]
    assert len({tuple(h[0]) for h in hyps}) > 10       # not a degenerate decode
    del m16
    m32 = _gpu_model(opt, sd, "fp32")
    h32, s32 = tr.translate_batch([m32], {"feats": feats})
    assert h32[:2048] == h32[2048:]
    pick = [0, 1, 511, 1024, 2047, 2048 + 3, 3000, 4095]
    sub = [f[pick].contiguous() for f in feats]
    h_sub, s_sub = tr.translate_batch([m32], {"feats": sub})
    assert h_sub == [h32[i] for i in pick]
    o_h, o_s, margins, _ = _oracle_margins(sd, opt, [f.cpu() for f in sub])
    for j, i in enumerate(pick):
        if h32[i] != o_h[j]:
            assert margins[j] < 1e-4, "video %d differs although the oracle margin is %g" % (i, margins[j])
        else:
            assert abs(s32[i][0] - o_s[j][0]) < 1e-4 * max(1.0, abs(o_s[j][0]))


def test_alternative_bos_token():
    """`ar_token_id` (Translator.py:33,61): the beams start from another token id than <bos>."""
    import care_b200
    rec = load_golden("cfg2_sharp")
    opt, sd, feats = rebuild_case(rec, batch=5)
    opt = dict(opt, ar_token_id=7)
    model = _gpu_model(opt, sd, "fp32")
    tr = care_b200.get_translator(opt)
    hyps, scores = tr.translate_batch([model], {"feats": [f.cuda() for f in feats]})
    o_h, o_s, margins, _ = _oracle_margins(sd, opt, feats)
    for v in range(5):
        assert hyps[v] == o_h[v] or margins[v] < 1e-4
    assert hyps != [h for h in rec["hyps"][:5]]     # really started from another token


@pytest.mark.parametrize("name", ["ens2_cfg2_sharp", "ens3_cfg2_k3"])
def test_model_ensembling_matches_reference_golden(name):
    """`translate_batch(models=[m1, m2, ...])`: beams follow the mean of the models' log-probabilities
    (models/Translator.py:39-52,111-133); fp32 mode against the unmodified reference's output."""
    import care_b200
    from tests.helpers import rebuild_ensemble_case
    rec = load_golden(name)
    opt, sds, feats = rebuild_ensemble_case(rec)
    models = [_gpu_model(opt, sd, "fp32") for sd in sds]
    tr = care_b200.get_translator(opt)
    hyps, scores = tr.translate_batch(models, {"feats": [f.cuda() for f in feats]})
    exact = sum(int(a == b) for a, b in zip(hyps, rec["hyps"]))
    assert exact == len(hyps), (hyps, rec["hyps"])
    for a, b in zip(scores, rec["scores"]):
        for x, y in zip(a, b):
            assert abs(x - y) < 1e-4 * max(1.0, abs(y))
    # 16-bit engines: same driver, well-formed output
    m16 = [_gpu_model(opt, sd, "fp16") for sd in sds]
    h16, _ = tr.translate_batch(m16, {"feats": [f.cuda() for f in feats]})
    assert len(h16) == len(hyps) and all(1 <= len(h[0]) <= opt["max_len"] - 1 for h in h16)
