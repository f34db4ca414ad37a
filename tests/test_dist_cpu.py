"""N>1 host logic on CPU: world_size-2 gloo process group -- shard ranges, weight broadcast, output gather."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cycle_diffusion_b200 import specs
from cycle_diffusion_b200.dist import shard_range
from tests.common import NARROW


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from cycle_diffusion_b200.dist import broadcast_state_dict, gather_images, shard_batch
    inv = [(n, s) for n, s, _ in specs.openai_unet_params(NARROW)]
    sd = specs.synth_state_dict(specs.openai_unet_params(NARROW), 11) if rank == 0 else None
    got = broadcast_state_dict(sd, inv, src=0)
    ref = specs.synth_state_dict(specs.openai_unet_params(NARROW), 11)
    same = all(torch.equal(got[k], ref[k]) for k in ref)
    batch = dict(image=torch.arange(7 * 3, dtype=torch.float32).view(7, 3), text=[f't{i}' for i in range(7)])
    mine = shard_batch(batch, rank, world)
    full = gather_images(mine['image'] * 2, dst=0)
    ok_gather = True if rank != 0 else torch.equal(full, batch['image'] * 2)
    q.put((rank, same, len(mine['text']), mine['text'][0], ok_gather))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    assert [shard_range(32, r, 8) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    assert [shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [shard_range(3, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 3), (3, 3)]
    cover = [i for r in range(5) for i in range(*shard_range(13, r, 5))]
    assert cover == list(range(13))


def test_two_rank_gloo_broadcast_shard_gather():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0] == (0, True, 4, 't0', True)
    assert res[1] == (1, True, 3, 't4', True)
