"""Multi-GPU plumbing.  The hot path shards by independent audio stream: stream s lives on rank
s % world_size for its whole life (its sliding windows, synthesizer ring and noise stream are sequential state),
so there is NO collective on the data path.  `torch.distributed` is used once, at init, to broadcast the model
weights from rank 0 (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
from typing import Dict, List, Optional

import numpy


def stream_assignment(n_streams: int, world_size: int) -> List[List[int]]:
    """streams handled by each rank (round-robin: stream s -> rank s % world_size)."""
    return [[s for s in range(n_streams) if s % world_size == r] for r in range(world_size)]


def broadcast_params(params: Optional[Dict[str, numpy.ndarray]], src: int = 0, device: str = 'cpu') -> Dict[str, numpy.ndarray]:
    """Every rank returns rank `src`'s dict of arrays (weights, statistics)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        meta[0] = [(k, tuple(v.shape), str(v.dtype)) for k, v in params.items()]
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for k, shape, dtype in meta[0]:
        if rank == src:
            t = torch.from_numpy(numpy.ascontiguousarray(params[k])).to(device)
        else:
            t = torch.empty(shape, dtype=getattr(torch, dtype), device=device)
        dist.broadcast(t, src=src)
        out[k] = t.cpu().numpy()
    return out
