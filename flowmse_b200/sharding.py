"""Utterance-level data parallelism for the sampler: one process per GPU, no collective on the data path.

The reference enhances files one by one on one GPU (/root/reference/evaluate.py:97-128) and has no multi-GPU
inference.  Utterances are independent, so the sharded path is: (1) rank 0 broadcasts the flat fp32 weight blob once,
(2) every rank runs the sampler on its own utterances, (3) the enhanced spectrograms are all-gathered at the end.
Assignment is length-aware LPT (cost ~ frames) so ranks finish together; utterances of equal padded length T are
batched together on a rank.  Works with any torch.distributed backend ("nccl" on GPUs; "gloo" in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def lpt_assign(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of utterance indices to ranks (cost = frames).
    Deterministic: ties broken by index, ranks by (load, rank)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * world_size
    out: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += int(lengths[i])
    for r in range(world_size):
        out[r].sort()
    return out


def bucket_by_length(indices: Sequence[int], lengths: Sequence[int], max_batch: int) -> List[List[int]]:
    """Group a rank's utterances into batches of identical padded length T (T % 64 == 0 gives few buckets)."""
    by_t: Dict[int, List[int]] = {}
    for i in indices:
        by_t.setdefault(int(lengths[i]), []).append(i)
    batches: List[List[int]] = []
    for t in sorted(by_t):
        idx = by_t[t]
        for k in range(0, len(idx), max_batch):
            batches.append(idx[k:k + max_batch])
    return batches


def broadcast_weights(blob: torch.Tensor, src: int = 0) -> torch.Tensor:
    """One broadcast of the packed fp32 backbone weights (250 MiB) from `src` to every rank."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def gather_ragged(local: List[Tuple[int, torch.Tensor]], n_total: int, max_T: int, device) -> List[torch.Tensor]:
    """All-gather per-utterance results of ragged length.

    `local` holds (utterance index, complex64 [1,F,T_i]) pairs of this rank.  Every rank contributes a zero-padded
    [n_local_max, F, max_T] block plus its index/length vectors; one all_gather each.  Returns the n_total results in
    utterance order on every rank.
    """
    world = dist.get_world_size() if dist.is_initialized() else 1
    F = local[0][1].shape[-2] if local else 256
    n_local = torch.tensor([len(local)], dtype=torch.int64, device=device)
    if world > 1:
        counts = [torch.zeros_like(n_local) for _ in range(world)]
        dist.all_gather(counts, n_local)
        n_max = int(max(c.item() for c in counts))
    else:
        n_max = len(local)
    meta = torch.full((n_max, 2), -1, dtype=torch.int64, device=device)
    block = torch.zeros((n_max, F, max_T, 2), dtype=torch.float32, device=device)
    for k, (idx, x) in enumerate(local):
        T = x.shape[-1]
        meta[k, 0], meta[k, 1] = idx, T
        block[k, :, :T] = torch.view_as_real(x.reshape(F, T))
    if world > 1:
        metas = [torch.empty_like(meta) for _ in range(world)]
        blocks = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(metas, meta)
        dist.all_gather(blocks, block)
    else:
        metas, blocks = [meta], [block]
    out: List[torch.Tensor] = [None] * n_total   # type: ignore
    for m, b in zip(metas, blocks):
        for k in range(m.shape[0]):
            idx, T = int(m[k, 0]), int(m[k, 1])
            if idx >= 0:
                out[idx] = torch.view_as_complex(b[k, :, :T].contiguous())[None]
    missing = [i for i, o in enumerate(out) if o is None]
    if missing:
        raise RuntimeError(f"gather_ragged: no rank produced utterances {missing[:5]}")
    return out


def enhance_sharded(specs: List[torch.Tensor], enhance_fn: Callable[[torch.Tensor], torch.Tensor], device,
                    max_batch: int = 16) -> List[torch.Tensor]:
    """Shard `specs` (complex64 [1,F,T_i], T_i % 64 == 0, identical list on every rank) across the process group, run
    `enhance_fn` (batched [B,1,F,T] -> [B,1,F,T]) on this rank's share and all-gather the results."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    lengths = [int(s.shape[-1]) for s in specs]
    mine = lpt_assign(lengths, world)[rank]
    local: List[Tuple[int, torch.Tensor]] = []
    for batch in bucket_by_length(mine, lengths, max_batch):
        Y = torch.stack([specs[i].reshape(1, specs[i].shape[-2], specs[i].shape[-1]) for i in batch]).to(device)
        X = enhance_fn(Y.contiguous())
        for k, i in enumerate(batch):
            local.append((i, X[k]))
    return gather_ragged(local, len(specs), max(lengths), device)
