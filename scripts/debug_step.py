"""Debug helper: compares encode outputs and step-1 intermediates of the CUDA path with torch fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
import care_b200
from oracle import care_oracle as co
from tests.helpers import load_golden, rebuild_case

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4_sharp"
precision = sys.argv[2] if len(sys.argv) > 2 else "fp32"
rec = load_golden(name)
opt, sd, feats = rebuild_case(rec, batch=3)
m = care_b200.get_framework(dict(opt, care_precision=precision)); m.load_state_dict(sd); m = m.eval().cuda()
eng = m.engine()
enc = m.encoding_phase([f.cuda() for f in feats])
torch.cuda.synchronize()
o = co.encoding_phase(sd, opt, feats)
def cmp(tag, a, b):
    a = a.float().cpu(); b = b.float().cpu()
    print("%-28s max|d|=%.3e  ref max=%.3e" % (tag, (a - b).abs().max().item(), b.abs().max().item()))
for k in ("encoder_hidden_states", "preds_attr", "semantic_hidden_states"):
    if k in o: cmp(k, enc[k], o[k])
if "semantic_labels" in o: print("labels equal", torch.equal(enc["semantic_labels"].cpu(), o["semantic_labels"]))
B, K = 3, opt["beam_size"]
trace = []
eng.ar_decode(enc, B, beam_size=K, topk=1, trace=trace, trace_logits=True, early_exit_every=0)
ws = {k[0]: v for k, v in eng._ws.items()}
d = opt["dim_hidden"]
# torch reference of step 29's last state is messy; redo step 1 explicitly
import ctypes
from care_b200._lib import check
kv = eng.cross_kv(enc["encoder_hidden_states"])
bufs, bst = eng._beam_buffers(B, K, K)
check(eng.lib.care_beam_init(eng.ctx, ctypes.byref(bst), 2, eng._stream()))
eng.decode_step(1, B, K, enc, kv, bufs, bst)
torch.cuda.synchronize()
g = o.get("semantic_hidden_states")
x = sd["decoder.embedding.word_embeddings.weight"][torch.full((B*K,), 2)] + sd["decoder.embedding.position_embeddings.weight"][0]
if g is not None: x = x + co.repeat_rows(g, K)
x0 = F.layer_norm(x, (d,), sd["decoder.embedding.LayerNorm.weight"], sd["decoder.embedding.LayerNorm.bias"], 1e-12)
cmp("x0", ws["x0"], x0)
L = "decoder.layers.0."
q = F.linear(x0, sd[L+"intra_attention.SDPA.query.weight"], sd[L+"intra_attention.SDPA.query.bias"])
k_ = F.linear(x0, sd[L+"intra_attention.SDPA.key.weight"], sd[L+"intra_attention.SDPA.key.bias"])
v_ = F.linear(x0, sd[L+"intra_attention.SDPA.value.weight"], sd[L+"intra_attention.SDPA.value.bias"])
cmp("qkv", ws["kv_cache"][0], torch.cat([q, k_, v_], 1))
x1 = F.layer_norm(F.linear(v_, sd[L+"intra_attention.dense.weight"], sd[L+"intra_attention.dense.bias"]) + x0, (d,), sd[L+"intra_attention.LayerNorm.weight"], sd[L+"intra_attention.LayerNorm.bias"], 1e-12)
cmp("x1", ws["x1"], x1)
mem = co.repeat_rows(o["encoder_hidden_states"], K)
inputs = {kk: co.repeat_rows(o[kk], K) for kk in co.decoder_input_keys(opt)}
x2 = co._attention(sd, "decoder.layers.0.inter_attention", opt, x1.unsqueeze(1), mem, torch.zeros(B*K, 1, mem.shape[1], dtype=torch.bool)).squeeze(1)
cmp("x2", ws["x2"], x2)
hh = torch.relu(F.linear(x2, sd[L+"ffn.dense1.weight"], sd[L+"ffn.dense1.bias"]))
cmp("ffn_h", ws["ffn_h"], hh)
x3 = F.layer_norm(F.linear(hh, sd[L+"ffn.dense2.weight"], sd[L+"ffn.dense2.bias"]) + x2, (d,), sd[L+"ffn.LayerNorm.weight"], sd[L+"ffn.LayerNorm.bias"], 1e-12)
cmp("x3", ws["x3"], x3)
lg = F.linear(x3, sd["cls_head.tgt_word_prj.weight"])
cmp("logits", ws["logits"][:, :opt["vocab_size"]], lg)
for st in trace[:6]:
    ids = []
    anc, hist = st["pre"]["anc"], st["pre"]["tok_hist"]; t = st["step"]
    rows = [[int(hist[v, p, int(anc[v, b, p])]) for p in range(t - 1)] + [int(hist[v, t - 1, b])] for v in range(B) for b in range(K)]
    ref = co.decoding_phase(sd, opt, torch.tensor(rows), inputs, last_time_step_logits=True)
    live = (st["pre"]["done"] == 0).repeat_interleave(K)
    if live.any(): cmp("step %d logits" % t, st["logits"][live], ref[live])
if "semantic_labels" in o:
    p = o["preds_attr"]; gl = enc["semantic_labels"].cpu(); ol = o["semantic_labels"]
    for v in range(B):
        if gl[v].tolist() != ol[v].tolist():
            idx = [i for i in range(30) if gl[v, i] != ol[v, i]]
            print("video", v, "ranks differing", idx, "oracle", ol[v, idx].tolist(), "gpu", gl[v, idx].tolist())
            print("  oracle probs at those ranks", p[v, ol[v, idx]].tolist(), " gpu preds", enc["preds_attr"].cpu()[v, gl[v, idx]].tolist())
