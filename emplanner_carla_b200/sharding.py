"""Multi-GPU plumbing for the batched planner: frames are independent, so the batch is sharded contiguously
across ranks (one process per GPU) and the only collective is one all-gather of the fixed-stride plan
records.  torch.distributed is plumbing only; works with NCCL (GPU) and gloo (CPU tests)."""
import torch
import torch.distributed as dist


def shard_bounds(total, world, rank):
    """Contiguous [lo, hi) frame range of `rank`; the first `total % world` ranks take one extra frame."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class GatherBuffer:
    """[world, B, stride] buffer whose slice [rank] is the rank's own output: kernels write straight into it,
    then one in-place all-gather fills the other slices (equal shard sizes required)."""

    def __init__(self, batch_per_rank, stride, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.full = torch.zeros((self.world, int(batch_per_rank), int(stride)), dtype=torch.float64, device=device)

    @property
    def mine(self):
        return self.full[self.rank]

    def all_gather(self):
        if self.world > 1:
            dist.all_gather_into_tensor(self.full.view(-1), self.mine.reshape(-1), group=self.group)
        return self.full.view(self.world * self.full.shape[1], self.full.shape[2])
