"""Multi-GPU plumbing: one process per GPU, images sharded, ONE collective (the weight-blob broadcast).

Replaces the reference's "8 unrelated single-GPU jobs over disjoint JSON ranges" (README.md:148-188, trainer.py:288-293)
and the per-process ``torch.load`` of the checkpoint (txt2img.py:25-42 / ddpm_ddim_wrapper.py:378-379).  The sampling path
has no cross-sample operation (GroupNorm and attention are per sample), so no per-step communication exists.
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous split of the batch dimension: rank r owns [lo, hi).  Remainders go to the lowest ranks."""
    assert 0 <= rank < world_size
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank, world_size):
    """Slice every [B, ...] tensor (or list of B strings) of a batch to this rank's shard; CFG pairs stay together."""
    out = {}
    for k, v in tensors.items():
        lo, hi = shard_range(len(v), rank, world_size)
        out[k] = v[lo:hi]
    return out


def broadcast_weights(nets, src=0):
    """Rank `src` has loaded + finalized `nets`; every other rank receives the packed blobs and adopts them.
    Works with the NCCL backend on the GPU box (blob_tensor() is a zero-copy view of engine memory)."""
    rank = dist.get_rank()
    for net in nets:
        dist.broadcast(net.blob_tensor(), src=src)
        if rank != src:
            net.adopt_blob()


def broadcast_state_dict(sd, keys_and_shapes, src=0):
    """Host-side variant (any backend, used by the gloo CPU tests): broadcast a reference-format state_dict tensor by
    tensor in inventory order.  Non-source ranks pass sd=None and receive a new dict."""
    rank = dist.get_rank()
    out = {}
    for name, shape in keys_and_shapes:
        t = sd[name].to(torch.float32).contiguous() if rank == src else torch.empty(shape, dtype=torch.float32)
        dist.broadcast(t, src=src)
        out[name] = t
    return out


def gather_images(img, dst=0):
    """Optional output gather to rank `dst` (mirrors distributed_concat, trainer.py:43-61).  Returns the full batch on
    `dst`, None elsewhere.  Shards may be ragged."""
    rank, world = dist.get_rank(), dist.get_world_size()
    sizes = [torch.zeros(1, dtype=torch.long, device=img.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([img.shape[0]], dtype=torch.long, device=img.device))
    mx = int(max(s.item() for s in sizes))
    pad = torch.zeros((mx,) + tuple(img.shape[1:]), dtype=img.dtype, device=img.device)
    pad[:img.shape[0]] = img
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    if rank != dst:
        return None
    return torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)])
