"""`Translator` plugins with the reference's names and signatures (models/Translator.py).

`get_translator(opt)` resolves `Translator_{decoding_type}`; `translate_batch(models, batch, vocab=,
teacher_model_wrapper=)` returns `(hyps, scores)` in the reference's formats.  The whole decode runs
on the device: beam state never visits the host until the final ids come back.
"""
import gc

import torch

from .engine import carry_n_best, ensemble_ar_decode, hyps_from_device

__all__ = ("Translator_ARFormer", "Translator_NARFormer", "get_translator")


def get_translator(opt: dict):
    # reference: models/Translator.py:14-19
    class_name = "Translator_{}".format(opt["decoding_type"])
    if class_name not in globals():
        raise ValueError("We can not find the class `{}` in {}".format(class_name, __file__))
    return globals()[class_name](opt)


def _single_model(models):
    if isinstance(models, (list, tuple)):
        if len(models) != 1:
            raise NotImplementedError("this entry point takes a single model (ensembles: Translator_ARFormer.translate_batch)")
        return models[0]
    return models


class Translator_ARFormer(object):
    """Beam search driver (reference: models/Translator.py:22-220)."""

    def __init__(self, opt: dict = {}):
        self.beam_size = opt.get("beam_size", 5)
        self.beam_alpha = opt.get("beam_alpha", 1.0)
        self.topk = opt.get("topk", 1)
        self.max_len = opt.get("max_len", 30)
        self.ar_token_id = opt.get("ar_token_id", None)   # alternative <bos> id (Translator.py:33,61)
        # translate_stream builds thousands of small Python lists per batch; every ~70k of them the interpreter runs a full
        # (generation-2) collection that walks every object of the process - 50-70 ms with a loaded model, six decodes of a
        # 512-video shard during which nothing is enqueued.  gc.freeze() at the start of a stream moves what is alive at
        # that point (model, workspaces, vocabulary) out of the collector's sight; opt["care_gc_freeze"] = False opts out.
        self.gc_freeze = bool(opt.get("care_gc_freeze", True))

    # host-resident batches larger than this are decoded in chunks whose host->device copies overlap
    # the previous chunk's decode (the copy of 4096 videos' fp32 features is 1.4 GB)
    pipeline_chunk = 2048

    def translate_batch(self, models, batch, *args, **kwargs):
        if isinstance(models, (list, tuple)) and len(models) > 1:
            return self._translate_ensemble(models, batch)
        model = _single_model(models)
        feats = batch["feats"]
        if feats[0].shape[0] == 0:
            return [], []
        with torch.no_grad():
            if feats[0].device.type == "cpu" and feats[0].shape[0] > self.pipeline_chunk:
                return self._translate_chunked(model, feats, self.pipeline_chunk)
            out = self.decode_on_device(model, feats)
        return hyps_from_device(*out, self.beam_alpha, self.topk)

    def _translate_chunked(self, model, host_feats, chunk):
        """One large host-resident batch as a stream of chunks: chunk i+1's host->device copy and chunk
        i-1's read-back + list building overlap chunk i's decode.  Videos are independent, so the result
        equals one big call; the reference's n_best carry-over (Translator.py:215) is re-applied across the
        chunk boundaries."""
        n_mod = len(model.engine().modality)
        host_feats = list(host_feats[:n_mod])
        B = host_feats[0].shape[0]
        # a short first chunk keeps the one copy nothing can hide (the first) small
        first = max(chunk // 4, 1)
        bounds = [(0, first)] + [(a, min(a + chunk, B)) for a in range(first, B, chunk)]
        chunks = ({"feats": [f[a:b] for f in host_feats]} for a, b in bounds)
        hyps, scores = [], []
        for h, s in self.translate_stream([model], chunks):
            hyps += h
            scores += s
        return carry_n_best(hyps, scores, self.topk)

    def _translate_ensemble(self, models, batch):
        """Model ensembling (reference: models/Translator.py:39-52,111-133): the beams follow the mean of the
        models' log-probabilities.  `batch['feats']` may hold one feature list per model (ModelEnsemble)."""
        feats = batch["feats"]
        per_model = isinstance(feats[0], (list, tuple))
        B = (feats[0][0] if per_model else feats[0]).shape[0]
        if B == 0:
            return [], []
        with torch.no_grad():
            encs = [m.encoding_phase(feats[i] if per_model else feats) for i, m in enumerate(models)]
            out = ensemble_ar_decode([m.engine() for m in models], encs, B, beam_size=self.beam_size, topk=self.topk,
                                     beam_alpha=self.beam_alpha, bos=self.ar_token_id)
        return hyps_from_device(*out, self.beam_alpha, self.topk)

    def translate_stream(self, models, batches, device_hook=None, **kwargs):
        """Throughput API for a stream of host-resident batches (the `for batch in loader` loop of
        translate.py:34-43): yields `(hyps, scores)` per batch, in order, with the same values as
        `translate_batch`.  While batch i decodes, the features of batch i+1 are copied host->device
        on a side stream, and the ids of batch i-1 (read back to pinned host memory behind their own
        decode) are turned into Python lists.
        `device_hook(out) -> out` (optional) runs on the device results before they are read back,
        e.g. the all-gather of care_b200.sharding for a video-sharded multi-GPU run."""
        model = _single_model(models)
        eng = model.engine()
        dev = eng.device
        n_mod = len(eng.modality)
        if self.gc_freeze:
            gc.freeze()
        main = torch.cuda.current_stream(dev)
        side = eng.copy_stream()
        staging = [None, None]
        landing = [None, None]      # pinned host copies of the results, two sets in flight
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def stage(i, batch):
            feats = list(batch["feats"][:n_mod])
            if feats[0].device.type != "cpu":
                return feats, None
            shapes = [tuple(f.shape) for f in feats]
            if staging[i % 2] is None or [tuple(t.shape) for t in staging[i % 2]] != shapes:
                # allocated in the copy stream's pool (a block recycled from the decode stream could still be
                # read by an earlier decode when the copy starts) and marked as used by the decode stream
                with torch.cuda.stream(side):
                    staging[i % 2] = [torch.empty(f.shape, dtype=f.dtype, device=dev) for f in feats]
                for t in staging[i % 2]:
                    t.record_stream(main)
            ev = torch.cuda.Event()
            with torch.cuda.stream(side):
                if i >= 2:
                    side.wait_event(freed[i % 2])
                for dst, src in zip(staging[i % 2], feats):
                    dst.copy_(src, non_blocking=True)
                ev.record(side)
            return staging[i % 2], ev

        it = iter(batches)
        try:
            nxt = stage(0, next(it))
        except StopIteration:
            return
        pending = None
        i = 0
        with torch.no_grad():
            while nxt is not None:
                cur = nxt
                try:
                    nxt = stage(i + 1, next(it))
                except StopIteration:
                    nxt = None
                if cur[1] is not None:
                    main.wait_event(cur[1])
                out = self.decode_on_device(model, cur[0], early_exit_every=0)
                freed[i % 2].record(main)
                if device_hook is not None:
                    out = device_hook(out)
                # the read-back of batch i is enqueued behind its decode and waited for one iteration
                # later, so the host builds batch i-1's lists while the device decodes batch i
                if landing[i % 2] is None or [tuple(t.shape) for t in landing[i % 2]] != [tuple(t.shape) for t in out]:
                    landing[i % 2] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in out]
                for dst, src in zip(landing[i % 2], out):
                    dst.copy_(src, non_blocking=True)
                landed = torch.cuda.Event()
                landed.record(main)
                if pending is not None:
                    pending[1].synchronize()
                    yield hyps_from_device(*pending[0], self.beam_alpha, self.topk)
                pending = (landing[i % 2], landed)
                i += 1
        if pending is not None:
            pending[1].synchronize()
            yield hyps_from_device(*pending[0], self.beam_alpha, self.topk)

    def decode_pipelined(self, model, host_feats, chunk):
        """Chunked decode of a host-resident batch: chunk i+1's features are copied on a side stream into
        the other staging set while chunk i decodes.  Videos are independent, so results are identical
        to one big call."""
        eng = model.engine()
        dev = eng.device
        n_mod = len(eng.modality)
        host_feats = list(host_feats[:n_mod])
        B = host_feats[0].shape[0]
        bounds = [(a, min(a + chunk, B)) for a in range(0, B, chunk)]
        main = torch.cuda.current_stream(dev)
        side = eng.copy_stream()
        staging = [[torch.empty((chunk,) + tuple(f.shape[1:]), dtype=f.dtype, device=dev) for f in host_feats]
                   for _ in range(2)]
        copied = [torch.cuda.Event() for _ in bounds]
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def enqueue_copy(i):
            a, b = bounds[i]
            with torch.cuda.stream(side):
                if i >= 2:
                    side.wait_event(freed[i % 2])
                for dst, src in zip(staging[i % 2], host_feats):
                    dst[:b - a].copy_(src[a:b], non_blocking=True)
                copied[i].record(side)

        enqueue_copy(0)
        outs = []
        for i, (a, b) in enumerate(bounds):
            if i + 1 < len(bounds):
                enqueue_copy(i + 1)
            main.wait_event(copied[i])
            outs.append(self.decode_on_device(model, [t[:b - a] for t in staging[i % 2]], early_exit_every=0))
            freed[i % 2].record(main)
        return tuple(torch.cat([o[j] for o in outs], dim=0) for j in range(4))

    def decode_on_device(self, model, feats, trace=None, early_exit_every=4):
        """Returns device tensors (ids [B, topk, T] int32 PAD-filled, lengths, raw scores, steps)."""
        eng = model.engine()
        if eng.max_len != self.max_len:
            raise ValueError("translator max_len %d != model max_len %d" % (self.max_len, eng.max_len))
        enc = model.encoding_phase(feats)
        B = enc["encoder_hidden_states"].shape[0]
        return eng.ar_decode(enc, B, beam_size=self.beam_size, topk=self.topk, beam_alpha=self.beam_alpha,
                             trace=trace, early_exit_every=early_exit_every,
                             bos=self.ar_token_id if self.ar_token_id is not None else None)


class Translator_NARFormer(object):
    """Length-beam + mask-predict driver (reference: models/Translator.py:223-318)."""

    def __init__(self, opt: dict = {}):
        self.opt = opt
        self.paradigm = opt.get("paradigm", "mp")
        if self.paradigm != "mp":
            raise NotImplementedError("only the mask-predict paradigm is on the accelerated hot path")
        self.max_len = opt["max_len"]
        self.length_beam_size = opt["length_beam_size"]
        self.beam_alpha = opt.get("beam_alpha", 1.0)
        self.length_bias = opt.get("length_bias", 0)

    def translate_batch(self, models, batch, teacher_model_wrapper=None, vocab=None):
        model = _single_model(models)
        if batch["feats"][0].shape[0] == 0:
            return [], []
        with torch.no_grad():
            teacher = self._teacher(model, teacher_model_wrapper, batch["feats"])
            tokens, lprobs = self.decode_on_device(model, batch["feats"], teacher=teacher)
        return tokens.cpu().tolist(), lprobs.cpu().tolist()

    def _teacher(self, model, teacher_model_wrapper, feats):
        """The auto-regressive teacher that rescores the candidates (reference: models/Translator.py:250-264;
        flags misc/Decoding/na_algorithms.py:29-32,98-104)."""
        if teacher_model_wrapper is None:
            self.vocab_mapping = None
            return None
        if not getattr(self, "_mapping_known", False):
            self.vocab_mapping = get_vocab_mapping(self.opt, teacher_model_wrapper.get_opt())
            self._mapping_known = True
        dev = model.engine().device
        teacher_model = teacher_model_wrapper.captioner.to(dev).eval()
        mapping = self.vocab_mapping.to(dev) if self.vocab_mapping is not None else None
        return dict(engine=teacher_model.engine(), enc=teacher_model.encoding_phase(feats), mapping=mapping,
                    masking=bool(self.opt.get("masking_decision", False)),
                    final=not self.opt.get("no_candidate_decision", False))

    def decode_on_device(self, model, feats, trace=None, teacher=None):
        """Returns device tensors: ids [B, 1, L] int32 (PAD after each caption's length) and per-token
        log-probabilities [B, 1, L] fp32, L = the longest length candidate of THIS batch (Translator.py:273)."""
        eng = model.engine()
        if eng.max_len != self.max_len:
            raise ValueError("translator max_len %d != model max_len %d" % (self.max_len, eng.max_len))
        enc = model.encoding_phase(feats)
        return eng.mask_predict(enc, self.opt, self.length_beam_size, self.length_bias, self.beam_alpha, trace=trace,
                                teacher=teacher)


def get_vocab_mapping(opt, teacher_opt):
    """student token id -> teacher token id when the two models were trained on different corpora (knowledge
    distillation changes the vocabulary), None when the vocabularies coincide (reference: models/Translator.py:321-344)."""
    import pickle
    if teacher_opt is None:
        return None
    with open(opt["info_corpus"], "rb") as f:
        vocab = pickle.load(f)["info"]["itow"]
    with open(teacher_opt["info_corpus"], "rb") as f:
        teacher_vocab = pickle.load(f)["info"]["itow"]
    if vocab == teacher_vocab:
        return None
    teacher_w2ix = {w: i for i, w in teacher_vocab.items()}
    mapping = torch.zeros(len(vocab), dtype=torch.long)
    for i, w in vocab.items():
        mapping[int(i)] = int(teacher_w2ix[w])
    assert int(mapping[0]) == 0   # <pad> stays <pad>
    return mapping
