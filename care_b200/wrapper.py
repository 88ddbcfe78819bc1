"""`Wrapper` API shell (reference: models/Wrapper.py, models/__init__.py) without pytorch-lightning.

Keeps what translate.py touches: `load_model_from_arguments` / `load_model` with the reference's defaults
(path rewriting, `modify_opt_if_necessary`, the strictness rules), `Model.load_from_checkpoint` on the
Lightning checkpoint layout (`state_dict` keys prefixed `captioner.`, `hyper_parameters['opt']`),
`ModelEnsemble`, `get_opt / get_vocab / get_references / get_keys_to_device / translate_step /
test_epoch_end / evaluation`, `.captioner`, `.translator`, `.eval_criterion`, `.eval() / .to()`.
"""
import itertools
import json
import os
import pickle
from collections import defaultdict
from types import SimpleNamespace
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from .criterion import get_criterion
from .framework import get_framework
from .translator import get_translator

PAD, EOS = 0, 3
BASE_DATA_PATH = "/data/video_datasets"   # reference: config/Constants.py:19


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


def to_sentence_with_tokenizer(hyp, tokenizer):
    """reference: misc/utils.py:140-149 (ids up to the first <eos>, decoded by an external tokenizer)."""
    end = hyp.index(EOS) if EOS in hyp else len(hyp)
    return tokenizer.decode(hyp[:end]).strip()


def save_dict_to_csv(path, file_name, dict_data):
    """reference: misc/utils.py:363-372: one row per call, header only when the file is created."""
    import pandas
    os.makedirs(path, exist_ok=True)
    if ".csv" not in file_name:
        file_name = file_name + ".csv"
    csv_path = os.path.join(path, file_name)
    exists = os.path.exists(csv_path)
    pandas.DataFrame([dict_data]).to_csv(csv_path, index=False, mode="a" if exists else "w", header=not exists)


def analyze_length_novel_unique(gt_data, data, vocab, splits, n=1):
    """reference: misc/utils.py:375-419: average caption length, share of captions never seen in the training
    split, share of distinct captions, number of distinct n-grams used."""
    grams, sents = set(), set()
    total_len = count = 0
    for items in data.values():
        for item in items:
            words = item["caption"].split(" ")
            sents.add(item["caption"])
            total_len += len(words)
            count += 1
            grams.update(" ".join(words[j:j + n]) for j in range(len(words) - n + 1))
    train_sents = set()
    for i in splits["train"]:
        for cap in gt_data["video%d" % int(i)]:
            train_sents.add(" ".join(vocab[wid] for wid in cap[1:-1]))
    novel = sum(1 for s in sents if s not in train_sents)
    return total_len / count, novel / count, len(sents) / count, len(grams)


def _default_scorer():
    """COCOScorer of the reference tree (misc/cocoeval.py wraps pycocoevalcap + Java).  Present when this package
    runs inside the reference's checkout with its requirements installed; None otherwise."""
    try:
        from misc.cocoeval import COCOScorer   # noqa: the reference's own module, not part of this package
        return COCOScorer()
    except Exception:
        return None


class ModelBase(nn.Module):
    def __init__(self, opt: Dict[str, Any], new_opt_used_to_override: Dict[str, Any] = {}):
        super().__init__()
        # reference: models/Wrapper.py:24-39 (`save_hyperparameters` -> self.hparams)
        self.hparams = SimpleNamespace(opt=dict(opt), new_opt_used_to_override=dict(new_opt_used_to_override))
        newest_opt = {**self.hparams.opt, **self.hparams.new_opt_used_to_override}
        self.captioner = get_framework(newest_opt)
        self.translator = get_translator(newest_opt)
        self.tokenizer = newest_opt.get("tokenizer", None)
        self.coco_eval = "lang" in newest_opt["crits"]
        self.eval_criterion = None
        self.logged = {}   # what the reference hands to Lightning's self.log / self.log_dict

    def log(self, name, value, **kwargs):
        self.logged[name] = value

    def log_dict(self, values, **kwargs):
        self.logged.update(values)

    # -- accessors (reference: models/Wrapper.py:296-309,393-409) ---------------------------------
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

    def get_references(self):
        if not hasattr(self, "references"):
            with open(self.hparams.opt["reference"], "rb") as f:
                self.references = pickle.load(f)
        return self.references

    def get_keys_to_device(self, *a, **k):
        if isinstance(self.captioner, list):
            keys = set()
            for captioner in self.captioner:
                keys |= set(captioner.get_keys_to_device(*a, **k))
            return list(keys)
        return self.captioner.get_keys_to_device(*a, **k)

    # -- steps (reference: models/Wrapper.py:41-73) -------------------------------------------------
    def validation_step(self, batch, batch_idx=None):
        if self.coco_eval:
            return self.translate_step(batch, vocab=self.get_vocab(), assert_only_a_caption_per_video=True)
        assert not isinstance(self.captioner, (tuple, list))
        assert self.eval_criterion is not None
        self.eval_criterion.get_loss({**self.captioner.feedforward_step(batch), **batch})

    def test_step(self, batch, batch_idx=None):
        return self.validation_step(batch, batch_idx)

    def validation_epoch_end(self, all_step_outputs, crit_prefix="vali", log_best=True):
        if self.coco_eval:
            return self.evaluation(all_step_outputs, references=self.get_references(), log_scores=True,
                                   log_best_ever_scores=True, crit_prefix=crit_prefix)
        loss_info = self.eval_criterion.get_loss_info()
        self.log_dict({"{}_{}".format(crit_prefix, k): v for k, v in loss_info.items()})
        if log_best and "mAP" in loss_info:
            if not hasattr(self, "best_mAP") or loss_info["mAP"] > self.best_mAP:
                self.best_mAP = loss_info["mAP"]
            self.log("best_mAP", self.best_mAP)
        self.eval_criterion.reset_loss_recorder()

    def forward(self, batch, **kwargs):
        vocab = kwargs.pop("vocab", None)
        return self.translate_step(batch, vocab=vocab if vocab is not None else self.get_vocab(), **kwargs)

    # -- the hot entry point (reference: models/Wrapper.py:158-212) ---------------------------------
    def translate_step(self, batch, vocab, assert_only_a_caption_per_video=False, verbose=False,
                       inference_latency=False):
        if hasattr(self, "preprocess_batch_before_translate_step"):
            self.preprocess_batch_before_translate_step(batch)
        models = self.captioner if isinstance(self.captioner, list) else [self.captioner]
        hyps_of_a_batch, scores_of_a_batch = self.translator.translate_batch(
            models=models, batch=batch, vocab=vocab, teacher_model_wrapper=getattr(self, "teacher_model_wrapper", None))
        if not inference_latency and len(models) == 1 and getattr(self, "eval_criterion", None) is not None:
            # the non-latency branch: teacher-forced pass -> concept mAP / length criteria (Wrapper.py:182-184)
            self.eval_criterion.get_loss({**self.captioner.feedforward_step(batch), **batch})
        preds = defaultdict(list)
        for i in range(len(hyps_of_a_batch)):
            video_id = batch["video_ids"][i]
            hyps, scores = hyps_of_a_batch[i], scores_of_a_batch[i]
            assert isinstance(hyps, list)
            if assert_only_a_caption_per_video:
                assert len(hyps) == 1
            for hyp, score in zip(hyps, scores):
                if self.tokenizer is None:
                    caption = to_sentence(hyp, vocab)
                else:
                    caption = to_sentence_with_tokenizer(hyp, self.tokenizer)
                if verbose:
                    print("{}: {}({})".format(video_id, caption, score))
                preds[video_id].append({"image_id": video_id, "caption": caption, "score": score})
        return preds

    # -- metrics plumbing (reference: models/Wrapper.py:75-149, 214-273) ----------------------------
    def test_epoch_end(self, all_step_outputs, log_scores=True, verbose=True, keys_added_to_scores=["seed"],
                       analyze=True, save_csv_path=None, scorer=None):
        if not self.coco_eval:
            return self.validation_epoch_end(all_step_outputs, "test", log_best=False)
        opt = self.get_opt()
        first = all_step_outputs[0]
        n_caption_per_video = len(first[next(iter(first))])
        if n_caption_per_video == 1:
            preds_for_completion = {}
            if opt.get("dataset") == "VATEX" and opt.get("feats", "") != "I3D":
                if opt.get("VATEX_I3D_preds_json", ""):
                    with open(opt["VATEX_I3D_preds_json"], "rb") as f:
                        preds_for_completion = json.load(f)
                else:
                    print("- Partial data is missing, only obtain the subset's performance")
            scores, detail_scores, pred_captions = self.evaluation(
                all_step_outputs, references=None, scorer=scorer, log_scores=log_scores, log_prefix="test",
                crit_prefix="test", preds_for_completion=preds_for_completion)
        else:
            print("- We do not run coco evaluation because each video has %d generated captions." % n_caption_per_video)
            scores, detail_scores, pred_captions = {}, None, {}
            for item in all_step_outputs:
                pred_captions.update(item)
        for key in keys_added_to_scores:
            value = opt[key]
            scores[key] = "-".join(str(x) for x in value) if isinstance(value, (tuple, list)) else value
        if analyze:
            info_corpus = self.get_info_corpus()
            ave_length, novel, unique, usage = analyze_length_novel_unique(
                info_corpus["captions"], pred_captions, vocab=self.get_vocab(), splits=info_corpus["info"]["split"], n=1)
            scores.update({"ave_length": ave_length, "novel": novel, "unique": unique, "usage": usage})
        if opt.get("save_csv", False):
            save_dict_to_csv(opt["checkpoint_path"] if save_csv_path is None else save_csv_path,
                             opt.get("csv_name", "test_result.csv"), scores)
        if opt.get("json_path", ""):
            assert "json_name" in opt.keys()
            os.makedirs(opt["json_path"], exist_ok=True)
            with open(os.path.join(opt["json_path"], opt["json_name"]), "w") as f:
                json.dump(pred_captions, f)
        if verbose:
            for k, v in scores.items():
                print(k + (": %s" % v if isinstance(v, str) else ": %g" % v))
        return scores, detail_scores, pred_captions

    def evaluation(self, all_step_outputs, references=None, scorer=None, log_scores=True, log_best_ever_scores=False,
                   log_prefix="", crit_prefix="", preds_for_completion={}):
        """Merges the per-step predictions, scores them with the COCO scorer when one is available (pycocoevalcap
        and Java are not part of this package: pass `scorer=` or run inside the reference's environment), adds the
        criterion table (mAP, F1) and resets the recorders."""
        preds = {}
        for item in all_step_outputs:
            preds.update(item)
        if len(preds_for_completion):
            missing = [k for k in preds_for_completion if k not in preds]
            for key in missing:
                preds[key] = preds_for_completion[key]
            print("- Adding %d missing predictions for evaluation" % len(missing))
        scorer = scorer if scorer is not None else _default_scorer()
        scores, detail_scores = {}, None
        if scorer is not None:
            references = references if references is not None else self.get_references()
            scores, detail_scores = scorer.score(references, preds, preds.keys())
            candidates = [scores["Bleu_4"], scores["METEOR"], scores["ROUGE_L"], scores["CIDEr"]]
            scores["Sum"] = sum(s for s, flag in zip(candidates, self.hparams.opt["metric_sum"]) if flag)
        else:
            print("- No COCO scorer available (pycocoevalcap / Java): caption metrics skipped")
        loss_info = None
        if getattr(self, "eval_criterion", None) is not None:
            loss_info = self.eval_criterion.get_loss_info()
            if "mAP" in loss_info:
                scores["mAP"] = loss_info.pop("mAP")
        if log_scores:
            self.log_dict({"{}_{}".format(log_prefix, k): v for k, v in scores.items()} if log_prefix else scores)
            if loss_info is not None:
                self.log_dict({"{}_{}".format(crit_prefix, k): v for k, v in loss_info.items()} if crit_prefix
                              else loss_info)
        if log_best_ever_scores and "Sum" in scores:
            if not hasattr(self, "best_Sum") or scores["Sum"] > self.best_Sum:
                self.best_Sum = scores["Sum"]
                self.CIDEr_in_the_best = scores["CIDEr"]
            if not hasattr(self, "best_CIDEr") or scores["CIDEr"] > self.best_CIDEr:
                self.best_CIDEr = scores["CIDEr"]
            self.log("best_Sum", self.best_Sum)
            self.log("best_CIDEr", self.best_CIDEr)
        if getattr(self, "eval_criterion", None) is not None:
            self.eval_criterion.reset_loss_recorder()
        return scores, detail_scores, preds

    # -- teacher for mask-predict rescoring (reference: models/Wrapper.py:275-300) -------------------
    def on_validation_epoch_start(self):
        self.prepare_auxiliary_info()

    def on_test_epoch_start(self):
        self.prepare_auxiliary_info()

    def on_validation_epoch_end(self):
        self.post_process_auxiliary_info()

    def on_test_epoch_end(self):
        self.post_process_auxiliary_info()

    def prepare_auxiliary_info(self):
        opt = self.get_opt()
        if opt["decoding_type"] == "NARFormer" and opt.get("teacher_path", "") \
                and not hasattr(self, "teacher_model_wrapper"):
            object.__setattr__(self, "teacher_model_wrapper", Model.load_from_checkpoint(opt["teacher_path"], strict=True))

    def post_process_auxiliary_info(self):
        if hasattr(self, "teacher_model_wrapper"):
            object.__delattr__(self, "teacher_model_wrapper")

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
    """reference: models/Wrapper.py:412-421 (the training criterion is training code and is not built)."""

    def __init__(self, opt, new_opt_used_to_override={}, merge_opt=False):
        if merge_opt:
            opt, new_opt_used_to_override = {**opt, **new_opt_used_to_override}, {}
        super().__init__(opt, new_opt_used_to_override)
        self.eval_criterion = get_criterion(self.get_opt(), skip_crit_list=["lang"], override_opt={"calculate_mAP": True})


class ModelEnsemble(ModelBase):
    """Several checkpoints decoded together: the translator averages the models' log-probabilities
    (reference: models/Wrapper.py:617-714).  Checkpoints may use different modalities: the merged opt asks
    the loader for the union and `preprocess_batch_before_translate_step` hands each model its own list."""

    def __init__(self, checkpoint_paths: List[str], new_opt_used_to_override: Dict[str, Any] = {},
                 map_location: Optional[torch.device] = None, strict: bool = True, WRAPPER=Model):
        assert isinstance(checkpoint_paths, list) and len(checkpoint_paths) >= 1
        captioners, modalities, opt = [], [], None
        for path in checkpoint_paths:
            model = WRAPPER.load_from_checkpoint(path, map_location="cpu" if map_location is None else map_location,
                                                 strict=strict)
            captioners.append(model.captioner)
            new_opt = model.hparams.opt
            modalities.append(new_opt["modality"])
            if opt is None:
                opt = new_opt
                continue
            for char in new_opt["modality"]:
                key = "feats_%s" % char
                if char in opt["modality"]:
                    # the same modality must come from the same feature files in every checkpoint
                    assert list(new_opt[key]) == list(opt[key]), "%s, %s" % (new_opt[key], opt[key])
                else:
                    opt[key] = new_opt[key]
        if len(set(modalities)) == 1:
            self_need_to_split = False
        else:
            opt["modality"] = "".join(dict.fromkeys("".join(modalities)))
            self_need_to_split = True
        super().__init__(opt, new_opt_used_to_override)
        del self.captioner
        # a plain list (not a ModuleList), as in the reference: `isinstance(self.captioner, list)` is how
        # translate_step and get_keys_to_device recognise an ensemble
        self.captioner = captioners
        self.need_to_split_feats = self_need_to_split
        self.modality_of_all_checkpoints = modalities

    def preprocess_batch_before_translate_step(self, batch):
        if self.need_to_split_feats and not isinstance(batch["feats"][0], (list, tuple)):
            order = self.hparams.opt["modality"]
            batch["feats"] = [[batch["feats"][order.index(char)] for char in modality]
                              for modality in self.modality_of_all_checkpoints]
        return batch

    def train(self, mode=True):
        for model in self.captioner:
            model.train(mode)
        return self

    def eval(self):
        for model in self.captioner:
            model.eval()
        return self

    def to(self, device):
        for model in self.captioner:
            model.to(device)
        return self

    def parameters(self, recurse=True):
        return itertools.chain(*[model.parameters() for model in self.captioner])


def modify_opt_if_necessary(args, model):
    """reference: models/__init__.py:7-32 (retrieval database variants select other feature files)."""
    opt = model.get_opt()
    datasets = getattr(args, "retrieval_datasets", None) or []
    if datasets:
        assert opt["feats_r"]
        assert "CLIP_ViT-B-32" in opt["feats_r"], opt["feats_r"]
        assert "unique" in opt["feats_r"], opt["feats_r"]
        folder = os.path.dirname(opt["feats_r"])
        if len(datasets) == 1 and datasets[0] == "MSRVTT":
            opt["feats_r"] = os.path.join(folder, "CLIP_ViT-B-32_unique.hdf5")
        else:
            opt["feats_r"] = os.path.join(folder, "CLIP_ViT-B-32_{}_unique.hdf5".format("-".join(datasets)))
    ratio = getattr(args, "retrieval_db_ratio", 100)
    if ratio is not None and ratio < 100:
        assert opt["feats_r"] or opt["feats_t"]
        suffix = "_ratio%.1f.hdf5" % ratio
        if opt["feats_r"]:
            if isinstance(opt["feats_r"], (list, tuple)):
                assert len(opt["feats_r"]) == 1
                opt["feats_r"] = opt["feats_r"][0]
            opt["feats_r"] = opt["feats_r"].replace(".hdf5", suffix)
            print("- Modify feats_r to", opt["feats_r"])
        if opt["feats_t"]:
            opt["feats_t"] = opt["feats_t"].replace(".hdf5", suffix)
            print("- Modify feats_t to", opt["feats_t"])
    model.hparams.opt = opt
    model.hparams.new_opt_used_to_override = {}
    return model


def load_model(checkpoint_path, new_opt_used_to_override={}, device=torch.device("cpu"), strict=True,
               WRAPPER=Model, replace_paths=True, base_data_path=None, ensemble_flag=None):
    """reference: models/__init__.py:92-152"""
    if ensemble_flag is None:
        ensemble_flag = isinstance(checkpoint_path, (list, tuple))
    if ensemble_flag:
        model = ModelEnsemble(list(checkpoint_path), new_opt_used_to_override=new_opt_used_to_override,
                              map_location="cpu", strict=strict, WRAPPER=WRAPPER)
    else:
        model = WRAPPER.load_from_checkpoint(checkpoint_path, new_opt_used_to_override=new_opt_used_to_override,
                                             map_location="cpu", strict=strict)
    if replace_paths:
        # released checkpoints carry their author's data paths; translate.py builds its loader from get_opt()
        opt = model.get_opt()
        ori = os.path.dirname(opt["info_corpus"])
        assert os.path.basename(ori) == opt["dataset"]
        ori = os.path.dirname(ori)
        now = base_data_path if base_data_path is not None else BASE_DATA_PATH

        def _replace(item):
            if isinstance(item, (list, tuple)):
                return [_replace(x) for x in item]
            assert type(item) is str
            return item.replace(ori, now)

        for key in ["feats_a", "feats_m", "feats_i", "feats_o", "feats_t", "feats_r", "reference", "info_corpus"]:
            if key in opt:
                opt[key] = _replace(opt[key])
        model.hparams.opt = opt
        model.hparams.new_opt_used_to_override = {}
    model.eval()
    model.to(device)
    return model


_WRAPPERS = {"Model": Model, "ModelEnsemble": ModelEnsemble}


def load_model_from_arguments(args, ignore_empty_attributes=[], replace_paths=True,
                              pluggin_func=modify_opt_if_necessary):
    """reference: models/__init__.py:35-89"""
    if getattr(args, "no_cuda", False) or getattr(args, "gpus", 1) == 0 or not torch.cuda.is_available():
        raise RuntimeError("care_b200 has no CPU path: a CUDA (sm_100a) device is required")
    device = torch.device("cuda")
    ensemble_flag = False
    if hasattr(args, "checkpoint_path"):
        assert type(args.checkpoint_path) is str
        path = args.checkpoint_path
    elif hasattr(args, "checkpoint_paths"):
        assert isinstance(args.checkpoint_paths, (list, tuple))
        path = args.checkpoint_paths
        if len(path) > 1:
            ensemble_flag = True
        else:
            path = path[0]
    else:
        raise AttributeError("Neither `checkpoint_path` or `checkpoint_paths` is found in the given arguments")
    wrapper = getattr(args, "wrapper", "Model")
    if wrapper not in _WRAPPERS:
        raise NotImplementedError("wrapper %r is training-side code outside the accelerated path" % wrapper)
    strict = False
    if getattr(args, "load_strictly", False) or getattr(args, "strict", False):
        strict = True
    elif hasattr(args, "with_backbones") and not args.with_backbones:
        strict = True
        del args.with_backbones
    base_data_path = getattr(args, "base_data_path", BASE_DATA_PATH)
    for attr in ignore_empty_attributes:
        if hasattr(args, attr) and not getattr(args, attr):
            delattr(args, attr)
    model = load_model(path, new_opt_used_to_override=vars(args), device=device, strict=strict,
                       WRAPPER=_WRAPPERS[wrapper], replace_paths=replace_paths, base_data_path=base_data_path,
                       ensemble_flag=ensemble_flag)
    if pluggin_func is not None:
        model = pluggin_func(args, model)
    return model
