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


class _CudaView:
    """Zero-copy torch view of a raw device pointer through __cuda_array_interface__."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class PeerGatherBuffer:
    """All-gather without a collective: every rank owns a [world, B, stride] buffer allocated by the library
    (emp_device_alloc), exports it with CUDA IPC and opens the other ranks' buffers.  The QP kernel then stores
    each finished plan record into all peers' buffers itself (emp_sl_plan_p2p_dev), so the gather overlaps the
    solves; `fence()` is the only cross-rank synchronisation (a 1-element all-reduce on the compute stream)."""

    def __init__(self, batch_per_rank, stride, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B, self.stride = int(batch_per_rank), int(stride)
        nbytes = self.world * self.B * self.stride * 8
        self.base = _lib.lib().emp_device_alloc(nbytes)
        if not self.base:
            raise _lib.EmpError("emp_device_alloc failed: %s" % _lib.lib().emp_last_error().decode())
        self.full = torch.as_tensor(_CudaView(self.base, (self.world, self.B, self.stride)), device=device)
        self._token = torch.zeros(1, device=device)
        self.peer_ptrs, self._opened = [], []
        if self.world > 1:
            handle = (C.c_ubyte * 64)()
            _lib.check(_lib.lib().emp_ipc_export(C.c_void_p(self.base), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            slice_bytes = self.rank * self.B * self.stride * 8          # where MY frames go in every buffer
            for r in range(self.world):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                _lib.check(_lib.lib().emp_ipc_open(hb, C.byref(p)))
                self._opened.append(p)
                self.peer_ptrs.append(p.value + slice_bytes)
        self.peer_array = (C.c_void_p * max(len(self.peer_ptrs), 1))(*self.peer_ptrs)

    @property
    def mine(self):
        return self.full[self.rank]

    @property
    def n_peers(self):
        return len(self.peer_ptrs)

    def fence(self):
        """Stream-ordered barrier: when it completes every rank's kernels (and their peer stores) have finished."""
        if self.world > 1:
            dist.all_reduce(self._token, group=self.group)

    def gathered(self):
        return self.full.view(self.world * self.B, self.stride)

    def close(self):
        for p in self._opened:
            self._lib.lib().emp_ipc_close(p)
        self._opened = []
        if self.base:
            self.full = None
            self._lib.lib().emp_device_free(self._C.c_void_p(self.base))
            self.base = None
