"""Video-sharded multi-GPU decode.

Videos are independent units (the reference keeps one `Beam` per video, models/Translator.py:59-64),
so rank g decodes the contiguous shard [g*B/N, (g+1)*B/N) with replicated weights and no data-path
collective.  The only exchange is ONE all-gather of the decoded ids at the end: per video an int32
record [T ids (PAD filled) | length | steps | score bits], ~0.5 MB for 4096 videos.

`torch.distributed` is plumbing: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n units: the first n % world ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict, rank: int, world: int) -> Dict:
    """This rank's slice of a reference-style batch dict (`feats` list, `video_ids`, per-video tensors)."""
    n = batch["feats"][0].shape[0]
    lo, hi = shard_range(n, rank, world)
    out = {}
    for k, v in batch.items():
        if k == "feats":
            out[k] = [f[lo:hi] for f in v]
        elif isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n:
            out[k] = v[lo:hi]
        elif isinstance(v, (list, tuple)) and len(v) == n:
            out[k] = list(v[lo:hi])
        else:
            out[k] = v
    return out


def pack_hypotheses(out_tok: torch.Tensor, out_len: torch.Tensor, out_score: torch.Tensor,
                    out_t: torch.Tensor) -> torch.Tensor:
    """[B, n_best, T] ids + [B, n_best] len/score/steps -> int32 [B, n_best, T + 3] (score as raw bits)."""
    return torch.cat([out_tok, out_len.unsqueeze(-1), out_t.unsqueeze(-1),
                      out_score.contiguous().view(torch.int32).unsqueeze(-1)], dim=-1).contiguous()


def unpack_hypotheses(payload: torch.Tensor):
    T = payload.shape[-1] - 3
    out_tok = payload[..., :T].contiguous()
    out_len = payload[..., T].contiguous()
    out_t = payload[..., T + 1].contiguous()
    out_score = payload[..., T + 2].contiguous().view(torch.float32)
    return out_tok, out_len, out_score, out_t


def gather_hypotheses(payload: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gathers the per-rank records into input order.  `payload`: this rank's [b_local, n_best, W]
    int32 records; returns [n_total, n_best, W] on every rank.  Shards may differ by one video, so the
    send buffer is padded to the largest shard."""
    world = dist.get_world_size(group)
    if world == 1:
        return payload
    per = (n_total + world - 1) // world
    send = payload.new_zeros((per,) + tuple(payload.shape[1:]))
    send[:payload.shape[0]] = payload
    recv = payload.new_empty((world * per,) + tuple(payload.shape[1:]))
    dist.all_gather_into_tensor(recv, send, group=group)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_total, r, world)
        parts.append(recv[r * per: r * per + (hi - lo)])
    return torch.cat(parts, dim=0)


def pack_nar(tokens: torch.Tensor, lprobs: torch.Tensor, max_len: int) -> torch.Tensor:
    """Mask-predict results of one rank -> int32 records [B, 1, 2 * max_len + 1]: ids (PAD = 0 beyond the rank's
    canvas length L), log-probability bits (0.0 = log 1 beyond L, what the reference holds at <pad> slots,
    na_algorithms.py:78-79) and L itself - ranks see different longest length candidates (Translator.py:273)."""
    B, _, L = tokens.shape
    rec = tokens.new_zeros((B, 1, 2 * max_len + 1), dtype=torch.int32)
    rec[..., :L] = tokens
    rec[..., max_len:max_len + L] = lprobs.contiguous().view(torch.int32)
    rec[..., 2 * max_len] = L
    return rec


def unpack_nar(payload: torch.Tensor, max_len: int):
    """Inverse of pack_nar over the gathered records; the canvas length of the whole batch is the largest L."""
    if payload.shape[0] == 0:
        return payload[..., :0], payload[..., :0].view(torch.float32)
    L = int(payload[..., 2 * max_len].max().item())
    return payload[..., :L].contiguous(), payload[..., max_len:max_len + L].contiguous().view(torch.float32)


def translate_sharded(translator, model, batch: Dict, group=None) -> Tuple[List, List]:
    """`translate_batch` (Translator_ARFormer or Translator_NARFormer) over the whole batch with the videos
    sharded over the ranks of `group`; every rank returns the full (hyps, scores) in input order.  A rank whose
    shard is empty (fewer videos than ranks) skips the decode and contributes zero records to the gather."""
    from .engine import hyps_from_device
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_total = batch["feats"][0].shape[0]
    if n_total == 0:
        return [], []
    local = shard_batch(batch, rank, world)
    n_local = local["feats"][0].shape[0]
    dev = model.engine().device if model is not None else local["feats"][0].device   # (None: CPU stand-in tests)
    nar = hasattr(translator, "length_beam_size")
    Tm = translator.max_len - 1
    with torch.no_grad():
        if nar:
            if n_local:
                payload = pack_nar(*translator.decode_on_device(model, local["feats"]), translator.max_len)
            else:
                payload = torch.zeros((0, 1, 2 * translator.max_len + 1), dtype=torch.int32, device=dev)
        elif n_local:
            payload = pack_hypotheses(*translator.decode_on_device(model, local["feats"]))
        else:
            payload = torch.zeros((0, translator.topk, Tm + 3), dtype=torch.int32, device=dev)
    full = gather_hypotheses(payload, n_total, group)
    if nar:
        tokens, lprobs = unpack_nar(full, translator.max_len)
        return tokens.cpu().tolist(), lprobs.cpu().tolist()
    return hyps_from_device(*unpack_hypotheses(full), translator.beam_alpha, translator.topk)
