"""Multi-GPU plumbing: utterances are independent, so ranks shard the batch and the only
collective is ONE broadcast of the packed weight buffer."""
from __future__ import annotations

import torch


def shard_indices(n_items: int, rank: int, world: int, lengths=None):
    """Indices of the utterances rank ``rank`` processes.  With ``lengths`` the items are
    dealt longest-first round-robin so per-rank step counts match (Tacotron2 decode length
    follows the input length); otherwise a strided slice batch[rank::world]."""
    if lengths is None:
        return list(range(rank, n_items, world))
    order = sorted(range(n_items), key=lambda i: (-int(lengths[i]), i))
    return sorted(order[rank::world])


def broadcast_packed(model, src: int = 0):
    """Rank ``src`` packs its weights; every rank receives the flat buffer with a single
    torch.distributed.broadcast (NCCL over NVLink on the GPU box, gloo in CPU tests) and
    rebuilds its pointer table from the (config-determined) layout."""
    import torch.distributed as dist
    packed = model.packed() if dist.get_rank() == src else model.empty_packed()
    dist.broadcast(packed.flat, src=src)
    model.use_packed(packed)
    return packed


def gather_audio(audio: torch.Tensor, dst: int = 0):
    """Optional end-of-run gather of per-rank waveforms (B_r, T) onto ``dst``."""
    import torch.distributed as dist
    world = dist.get_world_size()
    out = [torch.empty_like(audio) for _ in range(world)] if dist.get_rank() == dst else None
    dist.gather(audio, out, dst=dst)
    return out
