"""Multi-GPU plumbing of the hot path: streams shard one-per-rank (no data-path collective); the only exchange is the
shared-map gather of the fixed-capacity keypoint / descriptor slabs.  Backend-agnostic (NCCL on the GPU box,
gloo in the CPU tests).

A rank serves its streams as a few stream GROUPS (own extractor / matcher handles and CUDA stream each).  Every group owns packed
slabs -- keypoints | descriptors | counts | monoIndex in ONE contiguous buffer -- so that the exchange is ONE collective per group
and round, issued right behind that group's kernels on the group's own communicator: no rank-wide join, and with two slab sets
alternating per round the next round's kernels never wait for the network."""
import torch


def shard_streams(n_streams, rank, world):
    """Contiguous block of stream ids owned by `rank` (weak scaling: the caller usually fixes streams per rank instead)."""
    per = (n_streams + world - 1) // world
    lo = min(n_streams, rank * per)
    return list(range(lo, min(n_streams, lo + per)))


class PackedSlab:
    """Fixed-capacity output slabs of `nb` streams in one contiguous byte buffer: kps [nb, cap, 7] f32 (28-byte cv::KeyPoint rows),
    desc [nb, cap, 32] u8, n [nb] i32, mono [nb] i32 -- all views of `buf`."""

    def __init__(self, nb, cap, device):
        self.nb, self.cap = nb, cap
        o_desc = nb * cap * 28
        o_n = o_desc + nb * cap * 32
        o_mono = o_n + 4 * nb
        self.nbytes = (o_mono + 4 * nb + 15) // 16 * 16
        self.buf = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.kps = self.buf[:o_desc].view(torch.float32).view(nb, cap, 7)
        self.desc = self.buf[o_desc:o_n].view(nb, cap, 32)
        self.n = self.buf[o_n:o_mono].view(torch.int32)
        self.mono = self.buf[o_mono:o_mono + 4 * nb].view(torch.int32)

    @staticmethod
    def views_of(flat, nb, cap):
        """(kps, desc, n, mono) views of one rank's packed bytes inside a gathered buffer."""
        o_desc = nb * cap * 28
        o_n = o_desc + nb * cap * 32
        o_mono = o_n + 4 * nb
        return (flat[:o_desc].view(torch.float32).view(nb, cap, 7), flat[o_desc:o_n].view(nb, cap, 32), flat[o_n:o_mono].view(torch.int32),
                flat[o_mono:o_mono + 4 * nb].view(torch.int32))


class GroupSlabGather:
    """One all-gather per (group, round): out[g][d] is [world, slab bytes], byte-identical to the concatenation of the ranks' packed
    slabs.  Each group has its own communicator so that the groups' collectives do not serialise behind each other."""

    def __init__(self, dist, world, slabs):
        """slabs[g][d]: PackedSlab of group g, buffer set d."""
        self.dist, self.world = dist, world
        self.groups = [dist.new_group(list(range(world))) for _ in slabs]
        self.out = [[torch.empty((world, s.nbytes), dtype=torch.uint8, device=s.buf.device) for s in pair] for pair in slabs]
        self.slabs = slabs
        self.work = [[None for _ in pair] for pair in slabs]

    def wait(self, g, d):
        """Make the current stream wait for the last gather that read slabs[g][d] (before that set is overwritten)."""
        w = self.work[g][d]
        if w is not None:
            w.wait()
            self.work[g][d] = None

    def gather(self, g, d):
        """Enqueue the gather of slabs[g][d] behind the work already queued on the current stream; returns immediately."""
        self.work[g][d] = self.dist.all_gather_into_tensor(self.out[g][d].view(-1), self.slabs[g][d].buf, group=self.groups[g], async_op=True)
        return self.out[g][d]

    def drain(self):
        for g, pair in enumerate(self.work):
            for d in range(len(pair)):
                self.wait(g, d)


def max_over_ranks(dist, value, device):
    """Timing rule: every multi-GPU number is the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
