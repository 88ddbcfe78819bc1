"""Host-side logic of the multi-GPU path on CPU: shard arithmetic, the record format of the one
all-gather, and a world_size-2 gloo run of gather_hypotheses / translate_sharded with a stand-in
decoder (no GPU compute here; the real decode is covered by the -m gpu tests)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from care_b200 import sharding


def test_shard_range_covers_everything():
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_roundtrip():
    g = torch.Generator().manual_seed(0)
    tok = torch.randint(0, 1000, (5, 2, 29), generator=g, dtype=torch.int32)
    ln = torch.randint(1, 29, (5, 2), generator=g, dtype=torch.int32)
    t = ln.clone()
    sc = -torch.rand(5, 2, generator=g) * 50
    out = sharding.unpack_hypotheses(sharding.pack_hypotheses(tok, ln, sc, t))
    assert torch.equal(out[0], tok) and torch.equal(out[1], ln) and torch.equal(out[3], t)
    assert torch.equal(out[2], sc)   # bit pattern preserved


def test_shard_batch_slices_every_per_video_field():
    feats = [torch.arange(10 * 3 * 2).float().view(10, 3, 2), torch.arange(10 * 4).float().view(10, 4, 1)]
    batch = {"feats": feats, "video_ids": ["v%d" % i for i in range(10)], "category": torch.arange(10), "flag": 7}
    parts = [sharding.shard_batch(batch, r, 3) for r in range(3)]
    assert sum(len(p["video_ids"]) for p in parts) == 10
    assert torch.equal(torch.cat([p["feats"][0] for p in parts]), feats[0])
    assert torch.equal(torch.cat([p["category"] for p in parts]), batch["category"])
    assert all(p["flag"] == 7 for p in parts)


class _FakeTranslator:
    """Stands in for Translator_ARFormer: 'decodes' video v to tokens derived from its features."""
    beam_alpha, topk, max_len = 1.0, 1, 7

    def decode_on_device(self, model, feats):
        x = feats[0]
        B, T = x.shape[0], 6
        key = x.view(B, -1)[:, 0].to(torch.int32)
        tok = (key.view(B, 1, 1) + torch.arange(T, dtype=torch.int32).view(1, 1, T)).contiguous()
        ln = (key % 5 + 1).view(B, 1)
        tok = torch.where(torch.arange(T).view(1, 1, T) < ln.view(B, 1, 1), tok, torch.zeros_like(tok))
        score = -(key.float() + 0.5).view(B, 1)
        return tok, ln.to(torch.int32), score, ln.to(torch.int32)


class _FakeNarTranslator:
    """Stands in for Translator_NARFormer: the canvas length L depends on the shard's content."""
    max_len, length_beam_size = 9, 6

    def decode_on_device(self, model, feats):
        key = feats[0].view(feats[0].shape[0], -1)[:, 0].to(torch.int32)
        B = key.shape[0]
        lens = key % 5 + 4
        L = int(lens.max())
        pos = torch.arange(L, dtype=torch.int32).view(1, 1, L)
        tok = torch.where(pos < lens.view(B, 1, 1), key.view(B, 1, 1) + pos + 6, torch.zeros_like(pos))
        lp = torch.where(pos < lens.view(B, 1, 1), -(key.view(B, 1, 1) + pos).float() / 7, torch.zeros(1, 1, L))
        return tok.to(torch.int32).contiguous(), lp.contiguous()


def test_pack_unpack_nar_roundtrip():
    tr = _FakeNarTranslator()
    tok, lp = tr.decode_on_device(None, [torch.arange(6).float().view(6, 1, 1) + 3])
    rec = sharding.pack_nar(tok, lp, tr.max_len)
    assert rec.shape == (6, 1, 2 * tr.max_len + 1)
    t2, l2 = sharding.unpack_nar(rec, tr.max_len)
    assert torch.equal(t2, tok) and torch.equal(l2, lp)


def _worker(rank, world, port, n, q, nar=False):
    if nar:
        return _worker_nar(rank, world, port, n, q)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feats = [torch.arange(n).float().view(n, 1, 1) + 3]
        batch = {"feats": feats, "video_ids": list(range(n))}
        hyps, scores = sharding.translate_sharded(_FakeTranslator(), None, batch)
        q.put((rank, hyps, scores))
    finally:
        dist.destroy_process_group()


def _worker_nar(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        feats = [torch.arange(n).float().view(n, 1, 1) + 3]
        hyps, scores = sharding.translate_sharded(_FakeNarTranslator(), None, {"feats": feats})
        q.put((rank, hyps, scores))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1, 5, 8])
def test_translate_sharded_nar_gloo_world2(n):
    """Mask-predict results through the sharded path: ranks with different canvas lengths (and, for n = 1, an
    EMPTY shard on rank 1) still return the single-process answer on every rank."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q, True)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tok, lp = _FakeNarTranslator().decode_on_device(None, [torch.arange(n).float().view(n, 1, 1) + 3])
    for rank, hyps, scores in results:
        assert hyps == tok.tolist() and scores == lp.tolist(), rank


@pytest.mark.parametrize("n", [1, 7, 8])
def test_translate_sharded_gloo_world2(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process answer
    ref = _FakeTranslator().decode_on_device(None, [torch.arange(n).float().view(n, 1, 1) + 3])
    from care_b200.engine import hyps_from_device
    ref_h, ref_s = hyps_from_device(*ref, 1.0, 1)
    for rank, hyps, scores in results:
        assert hyps == ref_h and scores == ref_s, rank
