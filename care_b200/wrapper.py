"""`Wrapper` API shell (reference: models/Wrapper.py, models/__init__.py) without pytorch-lightning.

Keeps what translate.py touches: `load_model_from_arguments`, `Model.load_from_checkpoint` on the
Lightning checkpoint layout (`state_dict` keys prefixed `captioner.`, `hyper_parameters['opt']`),
`get_opt / get_vocab / get_keys_to_device / translate_step / test_epoch_end`, `.captioner`,
`.translator`, `.eval() / .to()`.
"""
import os
import pickle
from collections import defaultdict
from types import SimpleNamespace
from typing import Any, Dict, List

import torch
import torch.nn as nn

from .framework import get_framework
from .translator import get_translator

PAD, EOS = 0, 3


def to_sentence(hyp, vocab, break_words=(EOS, PAD), skip_words=(), extra_mappings={}, add_eos=False):
    """reference: misc/utils.py:117-137"""
    new_vocab = {**vocab, **extra_mappings} if len(extra_mappings) else vocab
    sent = []
    flag = False
    for word_id in hyp:
        if flag:
            break
        if word_id in skip_words:
            continue
        if word_id in break_words:
            if add_eos and word_id == EOS:
                flag = True
            else:
                break
        sent.append(new_vocab[word_id])
    return " ".join(sent)


class ModelBase(nn.Module):
    def __init__(self, opt: Dict[str, Any], new_opt_used_to_override: Dict[str, Any] = {}):
        super().__init__()
        # reference: models/Wrapper.py:24-39 (`save_hyperparameters` -> self.hparams)
        self.hparams = SimpleNamespace(opt=dict(opt), new_opt_used_to_override=dict(new_opt_used_to_override))
        newest_opt = {**self.hparams.opt, **self.hparams.new_opt_used_to_override}
        self.captioner = get_framework(newest_opt)
        self.translator = get_translator(newest_opt)
        self.tokenizer = newest_opt.get("tokenizer", None)
        if self.tokenizer is not None:
            raise NotImplementedError("external tokenizers are outside the accelerated hot path")
        self.coco_eval = "lang" in newest_opt["crits"]
        self.eval_criterion = None  # concept mAP criterion is eval-metric code (out of scope)

    # -- accessors (reference: models/Wrapper.py:296-309,393-403) ---------------------------------
    def get_opt(self):
        return {**self.hparams.opt, **self.hparams.new_opt_used_to_override}

    def update_opt(self, new_opt):
        self.hparams.opt.update(new_opt)

    def get_info_corpus(self):
        if not hasattr(self, "info_corpus"):
            with open(self.get_opt()["info_corpus"], "rb") as f:
                self.info_corpus = pickle.load(f)
        return self.info_corpus

    def get_vocab(self) -> Dict[int, str]:
        if getattr(self, "_vocab", None) is not None:
            return self._vocab
        return self.get_info_corpus()["info"]["itow"]

    def set_vocab(self, vocab: Dict[int, str]):
        """Synthetic runs have no corpus pickle; they install a vocab directly."""
        self._vocab = vocab

    def get_keys_to_device(self, *a, **k):
        return self.captioner.get_keys_to_device(*a, **k)

    def forward(self, batch, **kwargs):
        vocab = kwargs.pop("vocab", None)
        return self.translate_step(batch, vocab=vocab if vocab is not None else self.get_vocab(), **kwargs)

    # -- the hot entry point (reference: models/Wrapper.py:158-212) ---------------------------------
    def translate_step(self, batch, vocab, assert_only_a_caption_per_video=False, verbose=False,
                       inference_latency=False):
        hyps_of_a_batch, scores_of_a_batch = self.translator.translate_batch(
            models=[self.captioner], batch=batch, vocab=vocab, teacher_model_wrapper=None)
        preds = defaultdict(list)
        for i in range(len(hyps_of_a_batch)):
            video_id = batch["video_ids"][i]
            hyps, scores = hyps_of_a_batch[i], scores_of_a_batch[i]
            assert isinstance(hyps, list)
            if assert_only_a_caption_per_video:
                assert len(hyps) == 1
            for hyp, score in zip(hyps, scores):
                caption = to_sentence(hyp, vocab)
                if verbose:
                    print("{}: {}({})".format(video_id, caption, score))
                preds[video_id].append({"image_id": video_id, "caption": caption, "score": score})
        return preds

    def test_epoch_end(self, all_step_outputs, log_scores=True, verbose=False, save_csv_path="",
                       keys_added_to_scores=[], **kwargs):
        """reference: models/Wrapper.py:75-149.  Caption metrics need pycocoevalcap + Java (absent);
        predictions are merged and returned with empty score tables."""
        preds = {}
        for item in all_step_outputs:
            preds.update(item)
        return {}, {}, preds

    # -- checkpoint layout (reference: models/__init__.py:115-120,159-173; Wrapper.py:24-29) --------
    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, new_opt_used_to_override={}, map_location="cpu",
                             strict=True, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location, weights_only=False)
        hp = ckpt["hyper_parameters"]
        override = {**hp.get("new_opt_used_to_override", {}), **dict(new_opt_used_to_override)}
        model = cls(hp["opt"], override)
        sd = ckpt["state_dict"]
        captioner_sd = {k[len("captioner."):]: v for k, v in sd.items() if k.startswith("captioner.")}
        model.captioner.load_state_dict(captioner_sd, strict=strict)
        return model

    def to_checkpoint(self) -> Dict[str, Any]:
        """A dict in the Lightning layout the reference saves (train.py:76-96, save_weights_only)."""
        return {
            "state_dict": {"captioner." + k: v.detach().cpu() for k, v in self.captioner.state_dict().items()},
            "hyper_parameters": {"opt": dict(self.hparams.opt),
                                 "new_opt_used_to_override": dict(self.hparams.new_opt_used_to_override)},
        }


class Model(ModelBase):
    def __init__(self, opt, new_opt_used_to_override={}, merge_opt=False):
        if merge_opt:
            opt, new_opt_used_to_override = {**opt, **new_opt_used_to_override}, {}
        super().__init__(opt, new_opt_used_to_override)


def load_model(checkpoint_path, new_opt_used_to_override={}, device=torch.device("cpu"), strict=True,
               WRAPPER=Model, replace_paths=False, base_data_path=None, ensemble_flag=None):
    """reference: models/__init__.py:92-152 (single-model branch)."""
    if isinstance(checkpoint_path, (list, tuple)):
        if len(checkpoint_path) != 1:
            raise NotImplementedError("ModelEnsemble is outside the accelerated hot path")
        checkpoint_path = checkpoint_path[0]
    model = WRAPPER.load_from_checkpoint(checkpoint_path, new_opt_used_to_override=new_opt_used_to_override,
                                         map_location="cpu", strict=strict)
    if replace_paths:
        opt = model.get_opt()
        ori = os.path.dirname(os.path.dirname(opt["info_corpus"]))
        now = base_data_path if base_data_path is not None else ori

        def _replace(item):
            if isinstance(item, (list, tuple)):
                return [_replace(x) for x in item]
            return item.replace(ori, now)

        for key in ["feats_a", "feats_m", "feats_i", "feats_o", "feats_t", "feats_r", "reference", "info_corpus"]:
            if key in opt and opt[key]:
                opt[key] = _replace(opt[key])
        model.hparams.opt = opt
        model.hparams.new_opt_used_to_override = {}
    model.eval()
    model.to(device)
    return model


def load_model_from_arguments(args, ignore_empty_attributes=[], replace_paths=False, pluggin_func=None):
    """reference: models/__init__.py:35-89"""
    if getattr(args, "no_cuda", False) or getattr(args, "gpus", 1) == 0 or not torch.cuda.is_available():
        raise RuntimeError("care_b200 has no CPU path: a CUDA (sm_100a) device is required")
    device = torch.device("cuda")
    if hasattr(args, "checkpoint_path"):
        path = args.checkpoint_path
    elif hasattr(args, "checkpoint_paths"):
        path = args.checkpoint_paths
    else:
        raise AttributeError("Neither `checkpoint_path` or `checkpoint_paths` is found in the given arguments")
    strict = bool(getattr(args, "load_strictly", False) or getattr(args, "strict", False))
    for attr in ignore_empty_attributes:
        if hasattr(args, attr) and not getattr(args, attr):
            delattr(args, attr)
    model = load_model(path, new_opt_used_to_override=vars(args), device=device, strict=strict,
                       replace_paths=replace_paths, base_data_path=getattr(args, "base_data_path", None))
    if pluggin_func is not None:
        model = pluggin_func(args, model)
    return model
