"""Host-side plumbing of the Wrapper API shell (care_b200/wrapper.py) that needs no GPU: checkpoint loading with
the reference's defaults (models/__init__.py:35-152), ModelEnsemble's option merging (models/Wrapper.py:617-680),
test_epoch_end's score table / CSV / JSON (models/Wrapper.py:75-149) and the text utilities, differentially
against the reference's own functions where /root/reference is present."""
import json
import os
import pickle
from types import SimpleNamespace

import pytest
import torch

import care_b200
from care_b200 import wrapper as W
from synth.shapes import CONFIGS, make_opt


def _ckpt(tmp_path, name, opt):
    m = care_b200.Model(opt)
    path = os.path.join(str(tmp_path), name)
    torch.save(m.to_checkpoint(), path)
    return path


def _data_opt(base, cfg="cfg1", **over):
    return dict(make_opt(**CONFIGS[cfg]), dataset="MSVD", info_corpus=os.path.join(base, "MSVD", "info_corpus.pkl"),
                reference=os.path.join(base, "MSVD", "refs.pkl"), feats_m=[os.path.join(base, "MSVD", "feats", "m.hdf5")],
                feats_i=os.path.join(base, "MSVD", "feats", "i.hdf5"), feats_a=[], **over)


def test_load_model_rewrites_data_paths_by_default(tmp_path):
    path = _ckpt(tmp_path, "a.ckpt", _data_opt("/home/author/data"))
    model = care_b200.load_model(path, new_opt_used_to_override={"beam_size": 3})
    opt = model.get_opt()
    assert opt["info_corpus"] == "/data/video_datasets/MSVD/info_corpus.pkl"       # Constants.base_data_path
    assert opt["feats_m"] == ["/data/video_datasets/MSVD/feats/m.hdf5"] and opt["feats_a"] == []
    assert opt["feats_i"] == "/data/video_datasets/MSVD/feats/i.hdf5"
    assert opt["beam_size"] == 3 and model.hparams.new_opt_used_to_override == {}
    assert model.translator.beam_size == 3 and not model.training
    model = care_b200.load_model(path, base_data_path="/mnt/x")
    assert model.get_opt()["reference"] == "/mnt/x/MSVD/refs.pkl"
    model = care_b200.load_model(path, replace_paths=False)
    assert model.get_opt()["reference"] == "/home/author/data/MSVD/refs.pkl"
    with pytest.raises(AssertionError):      # the corpus folder must be named after the dataset (models/__init__.py:127)
        care_b200.load_model(_ckpt(tmp_path, "b.ckpt", dict(_data_opt("/x"), dataset="VATEX")))


def test_load_model_from_arguments_rules(tmp_path, monkeypatch):
    seen = {}

    def fake_load_model(checkpoint_path, **kw):
        seen.update(kw, checkpoint_path=checkpoint_path)
        return SimpleNamespace(get_opt=lambda: {"feats_r": "", "feats_t": ""}, hparams=SimpleNamespace())

    monkeypatch.setattr(W, "load_model", fake_load_model)
    with pytest.raises(RuntimeError):        # no CPU path
        W.load_model_from_arguments(SimpleNamespace(checkpoint_path="x.ckpt", no_cuda=True))
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    args = SimpleNamespace(checkpoint_path="x.ckpt", with_backbones=[], teacher_path="", beam_alpha=0.0, beam_size=4)
    W.load_model_from_arguments(args, ignore_empty_attributes=["teacher_path", "beam_alpha"])
    assert seen["strict"] is True and not hasattr(args, "with_backbones")       # models/__init__.py:66-68
    assert seen["replace_paths"] is True and seen["base_data_path"] == "/data/video_datasets"
    assert "teacher_path" not in seen["new_opt_used_to_override"] and seen["new_opt_used_to_override"]["beam_size"] == 4
    assert seen["ensemble_flag"] is False and seen["WRAPPER"] is care_b200.Model
    args = SimpleNamespace(checkpoint_paths=["a", "b"], with_backbones=["i"], base_data_path="/d")
    W.load_model_from_arguments(args)
    assert seen["strict"] is False and seen["ensemble_flag"] is True and seen["checkpoint_path"] == ["a", "b"]
    assert seen["base_data_path"] == "/d"
    W.load_model_from_arguments(SimpleNamespace(checkpoint_paths=["a"]))
    assert seen["checkpoint_path"] == "a" and seen["ensemble_flag"] is False
    with pytest.raises(AttributeError):
        W.load_model_from_arguments(SimpleNamespace())
    with pytest.raises(NotImplementedError):
        W.load_model_from_arguments(SimpleNamespace(checkpoint_path="x", wrapper="InterplayModel"))


def test_modify_opt_if_necessary():
    def model_with(opt):
        return SimpleNamespace(get_opt=lambda: dict(opt), hparams=SimpleNamespace(opt=None, new_opt_used_to_override={"a": 1}))

    opt = {"feats_r": "/d/MSRVTT/feats/CLIP_ViT-B-32_VATEX_unique.hdf5", "feats_t": ""}
    m = W.modify_opt_if_necessary(SimpleNamespace(retrieval_datasets=["MSRVTT"]), model_with(opt))
    assert m.hparams.opt["feats_r"] == "/d/MSRVTT/feats/CLIP_ViT-B-32_unique.hdf5" and m.hparams.new_opt_used_to_override == {}
    m = W.modify_opt_if_necessary(SimpleNamespace(retrieval_datasets=["MSRVTT", "VATEX"], retrieval_db_ratio=100), model_with(opt))
    assert m.hparams.opt["feats_r"] == "/d/MSRVTT/feats/CLIP_ViT-B-32_MSRVTT-VATEX_unique.hdf5"
    m = W.modify_opt_if_necessary(SimpleNamespace(retrieval_db_ratio=25.0),
                                  model_with({"feats_r": ["/d/r_unique.hdf5"], "feats_t": "/d/t.hdf5"}))
    assert m.hparams.opt["feats_r"] == "/d/r_unique_ratio25.0.hdf5" and m.hparams.opt["feats_t"] == "/d/t_ratio25.0.hdf5"


def test_model_ensemble_merges_modalities(tmp_path):
    base = "/home/author/data"
    a = _ckpt(tmp_path, "mi.ckpt", _data_opt(base, "cfg1"))
    opt_b = dict(_data_opt(base, "cfg1"), modality="ami", feats_a=[os.path.join(base, "MSVD", "feats", "a.hdf5")])
    b = _ckpt(tmp_path, "ami.ckpt", opt_b)
    ens = care_b200.load_model([a, b], new_opt_used_to_override={"beam_size": 2}, replace_paths=False)
    assert isinstance(ens.captioner, list) and len(ens.captioner) == 2 and ens.need_to_split_feats
    assert sorted(ens.hparams.opt["modality"]) == ["a", "i", "m"]
    assert ens.hparams.opt["feats_a"] == opt_b["feats_a"]
    assert ens.translator.beam_size == 2 and ens.eval_criterion is None
    order = ens.hparams.opt["modality"]
    batch = {"feats": [order.index(c) for c in order]}       # stand-ins: feature i "is" its index
    ens.preprocess_batch_before_translate_step(batch)
    assert batch["feats"] == [[order.index("m"), order.index("i")], [order.index("a"), order.index("m"), order.index("i")]]
    assert sorted(ens.get_keys_to_device()) == ["feats", "input_ids"]
    assert sum(p.numel() for p in ens.parameters()) == sum(p.numel() for c in ens.captioner for p in c.parameters())
    with pytest.raises(AssertionError):      # same modality, different feature files
        care_b200.load_model([a, _ckpt(tmp_path, "m2.ckpt", dict(_data_opt(base, "cfg1"), feats_m=["/other/m.hdf5"]))],
                             replace_paths=False)


class _FakeScorer:
    def score(self, references, preds, ids):
        assert set(ids) == set(preds)
        return {"Bleu_4": 0.4, "METEOR": 0.3, "ROUGE_L": 0.6, "CIDEr": 0.5}, {"per_video": len(preds)}


def test_test_epoch_end_tables_and_files(tmp_path):
    base = os.path.join(str(tmp_path), "data")
    os.makedirs(os.path.join(base, "MSVD"))
    vocab = {0: "<pad>", 1: "<unk>", 2: "<bos>", 3: "<eos>", 6: "a", 7: "man", 8: "runs", 9: "dog"}
    corpus = {"info": {"itow": vocab, "split": {"train": [0, 1]}},
              "captions": {"video0": [[2, 6, 7, 8, 3]], "video1": [[2, 6, 9, 8, 3], [2, 6, 7, 3]]}}
    with open(os.path.join(base, "MSVD", "info_corpus.pkl"), "wb") as f:
        pickle.dump(corpus, f)
    with open(os.path.join(base, "MSVD", "refs.pkl"), "wb") as f:
        pickle.dump({"video7": [{"image_id": "video7", "caption": "a man runs"}]}, f)
    out_dir = os.path.join(str(tmp_path), "out")
    opt = _data_opt(base, "cfg1", metric_sum=[1, 0, 1, 1], seed=7, save_csv=True, checkpoint_path=out_dir,
                    json_path=out_dir, json_name="preds.json", modality_list=["m", "i"])
    model = care_b200.Model(opt)
    steps = [{"video7": [{"image_id": "video7", "caption": "a man runs", "score": -0.5}]},
             {"video8": [{"image_id": "video8", "caption": "a dog dog", "score": -0.7}]}]
    scores, detail, preds = model.test_epoch_end(steps, verbose=False, keys_added_to_scores=["seed", "modality_list"],
                                                 scorer=_FakeScorer())
    assert scores["Sum"] == pytest.approx(0.4 + 0.6 + 0.5) and scores["seed"] == 7 and scores["modality_list"] == "m-i"
    assert scores["ave_length"] == 3.0 and scores["novel"] == 0.5 and scores["unique"] == 1.0 and scores["usage"] == 4
    assert detail == {"per_video": 2} and set(preds) == {"video7", "video8"}
    assert model.logged["test_CIDEr"] == 0.5
    assert json.load(open(os.path.join(out_dir, "preds.json"))) == preds
    model.test_epoch_end(steps, verbose=False, scorer=_FakeScorer())
    rows = open(os.path.join(out_dir, "test_result.csv")).read().strip().split("\n")
    assert len(rows) == 3 and rows[0].startswith("Bleu_4,")              # header once, one row per call
    # several captions per video: no COCO evaluation (Wrapper.py:104-110)
    multi = [{"v": [{"image_id": "v", "caption": "a man", "score": -1.0}, {"image_id": "v", "caption": "a dog", "score": -2.0}]}]
    scores, detail, preds = model.test_epoch_end(multi, verbose=False, analyze=False, keys_added_to_scores=[])
    assert detail is None and len(preds["v"]) == 2 and "Sum" not in scores


def test_text_utilities_against_the_reference():
    from oracle import ref_harness as rh
    vocab = {i: "w%d" % i for i in range(50)}
    hyps = [[7, 8, 9, 3, 5], [3], [4, 0, 6], [], [10, 11, 12]]
    assert [W.to_sentence(h, vocab) for h in hyps] == ["w7 w8 w9", "", "w4", "", "w10 w11 w12"]

    class Tok:
        def decode(self, ids):
            return " " + "".join(chr(97 + i) for i in ids) + " "

    assert W.to_sentence_with_tokenizer([1, 2, 3, 4], Tok()) == "bc" and W.to_sentence_with_tokenizer([5, 6], Tok()) == "fg"
    if not rh.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    rh.load_reference()
    from misc import utils as ru
    g = torch.Generator().manual_seed(3)
    for _ in range(50):
        h = torch.randint(0, 50, (int(torch.randint(0, 12, (1,), generator=g)),), generator=g).tolist()
        assert W.to_sentence(h, vocab) == ru.to_sentence(h, vocab)
        assert W.to_sentence_with_tokenizer(h, Tok()) == ru.to_sentence_with_tokenizer(h, Tok())
    gt = {"video%d" % i: [[2] + torch.randint(6, 12, (4,), generator=g).tolist() + [3] for _ in range(3)] for i in range(6)}
    data = {"video%d" % i: [{"caption": " ".join(vocab[t] for t in torch.randint(6, 12, (4,), generator=g).tolist())}]
            for i in range(40)}
    data["video1"] = [{"caption": " ".join(vocab[t] for t in gt["video2"][0][1:-1])}]
    splits = {"train": [0, 1, 2, 3]}
    assert W.analyze_length_novel_unique(gt, data, vocab, splits, n=1) == ru.analyze_length_novel_unique(gt, data, vocab, splits, n=1)
    assert W.analyze_length_novel_unique(gt, data, vocab, splits, n=2) == ru.analyze_length_novel_unique(gt, data, vocab, splits, n=2)
