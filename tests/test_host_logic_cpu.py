"""Host-side pieces of the reference-facing API that need no GPU: detokenisation (misc/utils.py:117-137),
the n_best carry-over of collect_hypothesis_and_scores (models/Translator.py:211-220), the plugin
registries and the options the accelerated path refuses."""
import pytest
import torch

import care_b200
from care_b200.engine import carry_n_best, hyps_from_device
from synth.shapes import CONFIGS, make_opt


def test_to_sentence_stops_at_eos_and_pad():
    vocab = {i: "w%d" % i for i in range(20)}
    assert care_b200.to_sentence([5, 6, 3, 7], vocab) == "w5 w6"
    assert care_b200.to_sentence([5, 0, 7], vocab) == "w5"
    assert care_b200.to_sentence([], vocab) == ""
    assert care_b200.to_sentence([5, 6, 3, 7], vocab, add_eos=True) == "w5 w6 w3"
    assert care_b200.to_sentence([5, 9, 6], vocab, skip_words=(9,)) == "w5 w6"


def test_to_sentence_matches_reference_when_present():
    from oracle import ref_harness as rh
    if not rh.reference_available():
        pytest.skip("reference tree not present (GPU box)")
    rh.load_reference()
    from misc.utils import to_sentence as ref_to_sentence
    g = torch.Generator().manual_seed(0)
    vocab = {i: "w%d" % i for i in range(50)}
    for _ in range(200):
        hyp = torch.randint(0, 50, (int(torch.randint(0, 12, (1,), generator=g)),), generator=g).tolist()
        assert care_b200.to_sentence(hyp, vocab) == ref_to_sentence(hyp, vocab)


def test_hyps_from_device_carries_n_best_over_like_the_reference():
    # video 0 has 3 hypotheses, video 1 only 1, video 2 has 3 again: the reference truncates video 2 to 1
    T = 6
    tok = torch.zeros(3, 3, T, dtype=torch.int32)
    ln = torch.tensor([[2, 3, 4], [2, 0, 0], [1, 2, 3]], dtype=torch.int32)
    for v in range(3):
        for r in range(3):
            tok[v, r, :ln[v, r]] = torch.arange(10 * v + r + 4, 10 * v + r + 4 + int(ln[v, r]))
    score = -torch.arange(9, dtype=torch.float32).view(3, 3) - 1.0
    hyps, scores = hyps_from_device(tok, ln, score, ln.clone(), 0.7, 3)
    assert [len(h) for h in hyps] == [3, 1, 1]
    assert hyps[0][1] == [5, 6, 7] and hyps[2][0] == [24]
    assert scores[0][2] == pytest.approx(-3.0 / 4 ** 0.7)


def test_chunked_lists_equal_one_call_after_carry_over():
    g = torch.Generator().manual_seed(5)
    for trial in range(20):
        B, topk, T = 23, 3, 7
        ln = torch.randint(0 if trial % 2 else 1, T, (B, topk), generator=g, dtype=torch.int32)
        ln[:, 0].clamp_(min=1)
        tok = torch.randint(4, 90, (B, topk, T), generator=g, dtype=torch.int32)
        sc = -torch.rand(B, topk, generator=g)
        tt = ln.clamp(min=1)
        whole = hyps_from_device(tok, ln, sc, tt, 0.7, topk)
        hyps, scores = [], []
        for a in range(0, B, 5):
            h, s = hyps_from_device(tok[a:a + 5], ln[a:a + 5], sc[a:a + 5], tt[a:a + 5], 0.7, topk)
            hyps += h
            scores += s
        assert carry_n_best(hyps, scores, topk) == whole


def test_registries_and_refusals():
    opt = make_opt(**CONFIGS["cfg2"])
    assert type(care_b200.get_translator(opt)).__name__ == "Translator_ARFormer"
    assert type(care_b200.get_translator(make_opt(**CONFIGS["cfg5"]))).__name__ == "Translator_NARFormer"
    with pytest.raises(ValueError):
        care_b200.get_translator(dict(opt, decoding_type="Nope"))
    for bad in (dict(decoder="LSTM_rnn"), dict(with_category=True), dict(transformer_pre_ln=True),
                dict(num_hidden_layers_decoder=2), dict(trainable_pe=False), dict(cls_head="MLPHead"),
                dict(use_attr_type="prefix"), dict(attr_layer_pos="parallel", use_attr_type="_att")):
        with pytest.raises(ValueError):
            care_b200.get_framework({**opt, **bad})
    m = care_b200.get_framework(opt)
    assert m.input_keys_for_decoder == ["encoder_hidden_states", "semantic_hidden_states"]
    assert care_b200.get_framework(make_opt(**CONFIGS["cab"])).input_keys_for_decoder == [
        "encoder_hidden_states", "semantic_embs"]
    with pytest.raises(RuntimeError):
        m.engine()            # parameters on the CPU: there is no CPU path
