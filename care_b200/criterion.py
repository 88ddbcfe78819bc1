"""Evaluation criteria of the non-latency eval branch (reference: misc/Crit/*).

`translate.py` without `--latency` runs, after every decoded batch, a teacher-forced `feedforward_step` and
feeds its outputs to `model.eval_criterion` (models/Wrapper.py:182-184, 420-421): for a CARE model that is the
concept-detection criterion `NoisyOrMIL` (misc/Crit/crit_attribute.py:14-109: normalised BCE, F1@{5..50} and
the mean average precision of the 500-way concept ranking), for NACF additionally the length criterion
(misc/Crit/crit_length.py).  Names, recorded fields and arithmetic follow the reference so that
`get_loss_info()` returns the same table; the per-video Python loop of the AP computation
(crit_attribute.py:76-91) is replaced by one batched expression on the tensors' own device (the ranking comes
from the same `sort` call, so ties among clamped probabilities fall as they do there).
"""
import copy
import json
import os
from typing import Dict, List

import torch


class AverageMeter(object):
    """reference: misc/logger.py:51-70"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1, multiply=True):
        self.val = val
        self.sum += val * n if multiply else val
        self.count += n
        self.avg = self.sum / self.count


class CritBase(object):
    """reference: misc/Crit/base.py:6-48"""

    def __init__(self, keys: List[str], weights=1.0, batch_mean: bool = True):
        self.keys = keys
        self.weights = weights
        self.batch_mean = batch_mean

    def _step(self, *inputs) -> torch.Tensor:
        raise NotImplementedError()

    def __call__(self, kwargs: Dict[str, torch.Tensor]):
        sources1, sources2, *others = [kwargs.get(key, None) for key in self.keys]
        if not isinstance(sources1, list):
            sources1 = [sources1]
        if not isinstance(sources2, list):
            sources2 = [sources2] * len(sources1)
        else:
            assert len(sources1) == len(sources2)
        if not isinstance(self.weights, list):
            self.weights = [self.weights] * len(sources1)
        assert len(sources1) == len(self.weights)
        loss = None
        dinominator = float(sources1[0].size(0)) if self.batch_mean else 1.0
        for i, (weight, src1, src2) in enumerate(zip(self.weights, sources1, sources2)):
            term = weight * self._step(i, src1, src2, *others) / dinominator
            loss = term if loss is None else loss + term
        return loss, dinominator


class NoisyOrMIL(CritBase):
    """reference: misc/Crit/crit_attribute.py:14-109 (flag 'V': the video-level concept head)."""

    def __init__(self, opt, keys=None):
        super().__init__(keys=["preds_attr", "avg_prob_attr", "labels_attr"] if keys is None else keys, batch_mean=True)
        self.topk_list = [5, 10, 20, 30, 40, 50]
        self.calculate_mAP = opt.get("calculate_mAP", False)
        if opt.get("attribute_prediction_sparse_sampling", False):
            raise NotImplementedError("attribute_prediction_sparse_sampling is a training-time regulariser "
                                      "(crit_attribute.py:22-23,52-58) outside the accelerated path")
        self.save_AP_path = opt.get("save_AP_path", None)
        if self.save_AP_path:
            self.all_AP = []

    def _step(self, index_indicator, preds_attr, avg_prob_attr, labels_attr, *others):
        assert not len(others)
        assert preds_attr.shape[1] <= labels_attr.shape[1]
        preds_attr = torch.clamp(preds_attr, 0.01, 0.99)
        labels_attr = labels_attr[:, :preds_attr.shape[1]].to(preds_attr.device)
        n_positive = labels_attr.sum(1).float()
        one = torch.tensor(1.0).to(preds_attr.device)
        loss = -(labels_attr * torch.log(preds_attr) + (1.0 - labels_attr) * torch.log(1.0 - preds_attr))
        loss = loss.sum(1) / torch.max(one, n_positive)
        if hasattr(self, "f1_recorder"):
            _, candidates = preds_attr.topk(max(self.topk_list), dim=1, sorted=True, largest=True)
            total_n_positive = labels_attr.sum(1)
            hits = labels_attr.gather(1, candidates).cumsum(1)
            sums = []
            for topk in self.topk_list:
                this_n_hit = hits[:, topk - 1].clone()
                this_n_hit[this_n_hit.eq(0)] = 1e-3
                precision = this_n_hit / topk
                recall = this_n_hit / total_n_positive
                f1 = 2 * precision * recall / (precision + recall)
                sums.append(f1.sum())
            for i, total in enumerate(torch.stack(sums).tolist()):     # one device read for the six F1 sums
                self.f1_recorder[i].update(total, labels_attr.size(0), multiply=False)
        if hasattr(self, "AP_recorder"):
            # AP of a video = mean over its positives of (positives ranked at or above it) / (its 1-based rank)
            _, indices = preds_attr.sort(dim=1, descending=True)
            lab_sorted = labels_attr.gather(1, indices)
            ranks = torch.arange(1, lab_sorted.shape[1] + 1, device=lab_sorted.device, dtype=torch.float32)
            precision = lab_sorted.cumsum(1) / ranks
            ap = (precision * lab_sorted).sum(1) / lab_sorted.sum(1)     # NaN without positives, as the reference's mean()
            for value in ap.tolist():
                self.AP_recorder.update(value, 1, multiply=False)
                if self.save_AP_path:
                    self.all_AP.append(value)
        return loss.sum()

    def get_fieldsnames(self, prefix=""):
        return ["%sF1-%02d" % (prefix, item) for item in self.topk_list] + \
            (["%smAP" % prefix] if hasattr(self, "AP_recorder") else [])

    def get_info(self):
        if self.save_AP_path:
            os.makedirs(os.path.dirname(self.save_AP_path), exist_ok=True)
            with open(self.save_AP_path, "w") as wf:
                json.dump(self.all_AP, wf)
        return self.get_fieldsnames(), [item.avg for item in self.f1_recorder] + \
            ([self.AP_recorder.avg] if hasattr(self, "AP_recorder") else [])

    def reset_recorder(self):
        self.f1_recorder = [AverageMeter() for _ in range(len(self.topk_list))]
        if self.calculate_mAP:
            self.AP_recorder = AverageMeter()


class KLDivLoss(CritBase):
    """reference: misc/Crit/crit_length.py:6-13"""

    def __init__(self, opt):
        super().__init__(keys=["preds_length", "length_target"], batch_mean=True)
        self.crit = torch.nn.KLDivLoss(reduction="none")

    def _step(self, index_indicator, preds_length, length_target, *others):
        return torch.sum(self.crit(preds_length, length_target.to(preds_length.device)))


class Criterion(object):
    """reference: misc/Crit/base.py:51-118"""

    def __init__(self, crit_objects, names, scales):
        assert len(crit_objects) == len(names) == len(scales)
        self.crit_objects = crit_objects
        self.num_loss = len(crit_objects)
        self.names = names
        self.scales = scales
        self.n_current_round = 0
        self.reset_loss_recorder()

    def set_scales(self, new_scales):
        assert len(new_scales) == len(self.scales)
        self.scales = new_scales

    def reset_loss_recorder(self):
        self.loss_recorder = [AverageMeter() for _ in range(self.num_loss)]
        for crit_object in self.crit_objects:
            if getattr(crit_object, "reset_recorder", None) is not None:
                crit_object.reset_recorder()

    def get_loss(self, results, **kwargs):
        loss = []
        for i in range(self.num_loss):
            i_loss, num_samples = self.crit_objects[i](results)
            loss.append(i_loss * self.scales[i])
            self.loss_recorder[i].update(i_loss.item(), num_samples)
        return torch.stack(loss, dim=0).sum(0)

    def get_loss_info(self):
        all_names = self.names.copy()
        all_info = [meter.avg for meter in self.loss_recorder]
        for crit_object in self.crit_objects:
            if getattr(crit_object, "get_info", None) is not None:
                this_name, this_info = crit_object.get_info()
                all_names += this_name
                all_info += this_info
        return {n: i for n, i in zip(all_names, all_info)}


def _crit_info_attribute(opt):
    """reference: misc/Crit/prepare.py:17-52 (flag 'V' only: the other flags score decoder-side embeddings of
    model variants outside the accelerated path)."""
    scales = opt.get("attribute_prediction_scales", 1.0)
    flags = opt["attribute_prediction_flags"]
    if not isinstance(scales, list):
        scales = [scales]
    elif len(scales) == 1:
        scales = scales * len(flags)
    else:
        assert len(scales) == len(flags), "#scales %d vs. #flags %d" % (len(scales), len(flags))
    objects, names = [], []
    for flag in flags:
        if flag != "V":
            raise NotImplementedError("attribute_prediction_flags %r: only the video-level concept head 'V' is on the "
                                      "accelerated path" % flags)
        names.append("%s-Attr" % flag)
        objects.append(NoisyOrMIL(opt))
    return objects, names, scales


def _crit_info_length(opt):
    """reference: misc/Crit/prepare.py:9-14"""
    return [KLDivLoss(opt)], ["Length Loss"], [opt.get("length_prediction_scale", 1.0)]


_CRIT_INFO = {"attribute": _crit_info_attribute, "length": _crit_info_length}


def get_criterion(opt, skip_crit_list=[], override_opt={}):
    """reference: misc/Crit/__init__.py:22-69.  The caption loss ('lang') is a training criterion: the
    evaluation criterion the wrapper builds always skips it (models/Wrapper.py:421)."""
    _opt = copy.deepcopy(opt)
    _opt.update(override_opt)
    assert isinstance(_opt["crits"], list)
    crit_objects, names, scales = [], [], []
    for crit in [item for item in _opt["crits"] if item not in skip_crit_list]:
        if crit not in _CRIT_INFO:
            raise NotImplementedError("criterion %r is training code outside the accelerated path" % crit)
        o, n, s = _CRIT_INFO[crit](_opt)
        assert len(o) == len(n) == len(s)
        crit_objects.extend(o)
        names.extend(n)
        scales.extend(s)
    if not len(crit_objects):
        return None
    return Criterion(crit_objects=crit_objects, names=names, scales=scales)
