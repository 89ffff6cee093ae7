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
    """All-gather without a collective and without a barrier.

    Every rank owns a [world, B, stride] buffer plus 2 x world flag words, allocated by the library
    (emp_device_alloc), exported with CUDA IPC and opened by the other ranks.  The QP kernel stores each finished plan
    record into all peers' buffers itself (NVLink P2P stores overlapping the solves still running) and its last warp
    publishes the launch's epoch to this rank's `ready` word in every buffer (st.release.sys).  Consumers acquire the
    flags (`consume()`: a 1-CTA kernel spinning on ld.acquire.sys) only when they read, and hand the buffer back to the
    producers through the `free` words (credits), which `plan_qp()` waits for before it overwrites the peers' memory
    again.  NCCL / gloo is used once, to exchange the IPC handles.

        ready[p] in rank c's memory: written by producer p   - "my shard of epoch e is complete in your buffer"
        free[c]  in rank p's memory: written by consumer c   - "I am done with epoch e of my buffer, overwrite it"
    """

    FLAG_TIMEOUT_MS = 2000

    def __init__(self, batch_per_rank, stride, device, group=None):
        import ctypes as C
        from . import _lib
        self._lib, self._C = _lib, C
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.B, self.stride = int(batch_per_rank), int(stride)
        data_bytes = self.world * self.B * self.stride * 8
        self._flag_off = (data_bytes + 255) // 256 * 256
        nbytes = self._flag_off + 2 * self.world * 8
        self.base = _lib.lib().emp_device_alloc(nbytes)                  # zero-filled
        if not self.base:
            raise _lib.EmpError("emp_device_alloc failed: %s" % _lib.lib().emp_last_error().decode())
        self.full = torch.as_tensor(_CudaView(self.base, (self.world, self.B, self.stride)), device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.epoch = 0
        torch.cuda.synchronize(device)
        self.peer_ptrs, self._opened = [], []
        ready_ptrs, free_ptrs = [], []
        if self.world > 1:
            handle = (C.c_ubyte * 64)()
            _lib.check(_lib.lib().emp_ipc_export(C.c_void_p(self.base), handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)      # also: every buffer is zeroed by now
            slice_bytes = self.rank * self.B * self.stride * 8          # where MY frames go in every buffer
            for r in range(self.world):
                if r == self.rank:
                    continue
                p = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
                _lib.check(_lib.lib().emp_ipc_open(hb, C.byref(p)))
                self._opened.append(p)
                self.peer_ptrs.append(p.value + slice_bytes)
                ready_ptrs.append(p.value + self._flag_off + 8 * self.rank)
                free_ptrs.append(p.value + self._flag_off + 8 * (self.world + self.rank))
        n = max(len(self.peer_ptrs), 1)
        self.peer_array = (C.c_void_p * n)(*self.peer_ptrs)
        self._ready_array = (C.c_void_p * n)(*ready_ptrs)
        self._ready_local = self.base + self._flag_off                  # [world] arrival flags of MY buffer
        self._free_local = self.base + self._flag_off + 8 * self.world  # [world] credits for MY writes
        self._self_ready = self._ready_local + 8 * self.rank
        free_ptrs.append(self._free_local + 8 * self.rank)              # my own credit word
        self._free_array = (C.c_void_p * len(free_ptrs))(*free_ptrs)

    @property
    def mine(self):
        return self.full[self.rank]

    @property
    def n_peers(self):
        return len(self.peer_ptrs)

    def _stream(self):
        return self._C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def plan_qp(self, scenes, qp, n_enr_max, order=None):
        """The corridor + QP launch of one planning step on torch's current stream, writing this rank's shard into
        every rank's buffer and publishing a new epoch.  (The DP launch that precedes it only writes `mine`.)"""
        L, C = self._lib.lib(), self._C
        self.epoch += 1
        if self.epoch > 1:                                              # credits: everybody consumed the previous epoch
            self._lib.check(L.emp_gather_wait_signal_dev(C.c_void_p(self._free_local), self.world, self.epoch - 1, None,
                                                         0, C.c_void_p(self.err.data_ptr()), self.FLAG_TIMEOUT_MS,
                                                         self._stream()))
        kmax = (scenes.shape[1] - 6) // 2
        self._lib.check(L.emp_sl_qp_gather_dev(C.c_void_p(scenes.data_ptr()), kmax, scenes.shape[0], C.byref(qp),
                                               int(n_enr_max), C.c_void_p(self.mine.data_ptr()), self.n_peers,
                                               self.peer_array, self.epoch, C.c_void_p(self._self_ready),
                                               self._ready_array, C.c_void_p(order.data_ptr()) if order is not None else None,
                                               self._stream()))

    def consume(self):
        """Enqueue the consumer side on torch's current stream: wait until every rank's shard of the current epoch
        has arrived in this buffer, then give the buffer back to all producers.  Work enqueued after this call on the
        same stream may read `gathered()`; it must be enqueued BEFORE the call if it has to finish before the
        buffer is overwritten - use `wait_arrival()` + `release()` around it in that case."""
        L, C = self._lib.lib(), self._C
        self._lib.check(L.emp_gather_wait_signal_dev(C.c_void_p(self._ready_local), self.world, self.epoch,
                                                     self._free_array, len(self._free_array),
                                                     C.c_void_p(self.err.data_ptr()), self.FLAG_TIMEOUT_MS, self._stream()))

    def wait_arrival(self):
        L, C = self._lib.lib(), self._C
        self._lib.check(L.emp_gather_wait_signal_dev(C.c_void_p(self._ready_local), self.world, self.epoch, None, 0,
                                                     C.c_void_p(self.err.data_ptr()), self.FLAG_TIMEOUT_MS, self._stream()))

    def release(self):
        L, C = self._lib.lib(), self._C
        self._lib.check(L.emp_gather_wait_signal_dev(None, 0, self.epoch, self._free_array, len(self._free_array),
                                                     C.c_void_p(self.err.data_ptr()), self.FLAG_TIMEOUT_MS, self._stream()))

    def check(self):
        """Host-side check after a synchronize: no flag wait timed out."""
        if int(self.err.item()) != 0:
            raise self._lib.EmpError("a gather flag wait timed out (a peer rank never delivered its shard)")

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
