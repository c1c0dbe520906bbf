"""N > 1 host logic on CPU: two gloo ranks, rank 0's weights broadcast to rank 1, streams sharded round-robin."""
import hashlib
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from realtime_yukarin_b200 import synthetic
from realtime_yukarin_b200.distributed import broadcast_params, stream_assignment


def _digest(d):
    h = hashlib.sha256()
    for k in sorted(d):
        h.update(k.encode())
        h.update(np.ascontiguousarray(d[k]).tobytes())
    return h.hexdigest()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    params = synthetic.make_unet_params(7, ndim=1, in_ch=9, out_ch=9, base=8) if rank == 0 else None
    got = broadcast_params(params, src=0, device='cpu')
    q.put((rank, _digest(got), len(got)))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_broadcast_two_ranks_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0][1] == results[1][1] and results[0][2] == results[1][2] > 30
    assert results[0][1] == _digest(synthetic.make_unet_params(7, ndim=1, in_ch=9, out_ch=9, base=8))


def test_stream_sharding_is_a_partition():
    for n, w in ((8, 8), (64, 8), (5, 2), (3, 4)):
        parts = stream_assignment(n, w)
        assert len(parts) == w
        flat = sorted(s for p in parts for s in p)
        assert flat == list(range(n))
        assert all(s % w == r for r, p in enumerate(parts) for s in p)
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
