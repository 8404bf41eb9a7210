"""Multi-GPU plumbing of the hot path: streams shard one-per-rank (no data-path collective); the only exchange is the
shared-map gather of the fixed-capacity keypoint / descriptor slabs.  Backend-agnostic (NCCL on the GPU box,
gloo in the CPU tests)."""
import torch


def shard_streams(n_streams, rank, world):
    """Contiguous block of stream ids owned by `rank` (weak scaling: the caller usually fixes streams per rank instead)."""
    per = (n_streams + world - 1) // world
    lo = min(n_streams, rank * per)
    return list(range(lo, min(n_streams, lo + per)))


class SlabGather:
    """One all-gather per step of (keypoints [B, cap, 7] f32-bytes, descriptors [B, cap, 32] u8, counts [B] i32).
    Slabs have fixed capacity, so the gathered buffers are byte-identical to the concatenation of the per-rank slabs."""

    def __init__(self, dist, world, kps, desc, n):
        self.dist, self.world = dist, world
        # concatenation along dim 0 (the form every backend accepts); returned as views [world, ...]
        self.flat = [torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for t in (kps, desc, n)]
        self.out = [f.view((world,) + tuple(t.shape)) for f, t in zip(self.flat, (kps, desc, n))]

    def __call__(self, kps, desc, n):
        for src, dst in zip((kps, desc, n), self.flat):
            self.dist.all_gather_into_tensor(dst, src.contiguous())
        return self.out


def max_over_ranks(dist, value, device):
    """Timing rule: every multi-GPU number is the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
