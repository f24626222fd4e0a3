"""Trajectory sharding across the GPUs of one box.

The batch shards by trajectory: rank r owns a contiguous range, solves it with no inter-GPU
traffic, and results are optionally collected on one rank with a single gather.  Works with
any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""


def shard_range(total, world_size, rank):
    """Contiguous range [start, start + count) of `total` trajectories owned by `rank`:
    ceil(total / world_size) per rank, the tail rank(s) take what is left."""
    per = -(-total // world_size)
    start = min(total, rank * per)
    return start, max(0, min(per, total - start))


def weak_scaling_seed_base(base_seed, per_gpu, rank):
    """Seed of the first trajectory of `rank` when every rank solves `per_gpu` problems: the
    global batch is seeds base_seed .. base_seed + world*per_gpu - 1, rank-major."""
    return base_seed + rank * per_gpu


def gather_to_rank0(local, total, dist, dst=0):
    """Collect the per-rank blocks (first dimension = trajectories of shard_range) on `dst`.
    Ranks may hold different counts; blocks are padded to the common size for the collective.
    Returns the [total, ...] tensor on dst, None elsewhere."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-total // world)
    padded = local
    if local.shape[0] < per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded = torch.cat([local, pad], dim=0)
    padded = padded.contiguous()
    if rank == dst:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, parts, dst=dst)
        out = torch.cat([parts[r][:shard_range(total, world, r)[1]] for r in range(world)], dim=0)
        return out
    dist.gather(padded, None, dst=dst)
    return None


class PeerGatherBuffer:
    """A device buffer on rank `dst` that every rank of the node can STORE into over NVLink / NVSwitch
    (CUDA IPC peer mapping), so that the solve kernel itself delivers its coefficient block to the
    collecting GPU: pass `view(rank)` as the `coeffs` argument of solve_standard and the kernel's
    32-byte stores go straight to the peer -- the gather of SURVEY.md section 8e without a separate
    collective and without an intermediate copy.  One process per GPU, single node.

    The buffer is one cudaMalloc on `dst` (not a slice of torch's caching allocator: an IPC handle
    names a whole allocation).  `finish()` is the only synchronisation a consumer needs: a barrier
    after every rank's stream has drained."""

    def __init__(self, dist, rows_per_rank, row_shape, dtype, dst=0):
        import numpy as np
        import torch
        from cuda.bindings import runtime as rt
        self._rt, self._torch, self._dist = rt, torch, dist
        self.rank, self.world, self.dst = dist.get_rank(), dist.get_world_size(), dst
        self.rows, self.row_shape, self.dtype = rows_per_rank, tuple(row_shape), dtype
        self.row_elems = int(np.prod(self.row_shape))
        self.itemsize = torch.empty((), dtype=dtype).element_size()
        self.bytes = self.world * rows_per_rank * self.row_elems * self.itemsize
        self._owner = self.rank == dst
        handle_bytes = [None]
        self.ptr = 0
        if self._owner:
            try:
                err, ptr = rt.cudaMalloc(self.bytes)
                self._check(err, "cudaMalloc")
                self.ptr = int(ptr)
                err, handle = rt.cudaIpcGetMemHandle(ptr)
                self._check(err, "cudaIpcGetMemHandle")
                handle_bytes = [bytes(handle.reserved)]
            except Exception as exc:          # every rank must still leave the broadcast below
                handle_bytes = [None]
                self._owner_error = exc
                if self.ptr:                  # the allocation succeeded but could not be exported
                    rt.cudaFree(self.ptr)
                    self.ptr = 0
        dist.broadcast_object_list(handle_bytes, src=dst)
        if handle_bytes[0] is None:
            raise RuntimeError("PeerGatherBuffer: rank %d could not export its buffer" % dst)
        ok, error = 1, None
        if not self._owner:
            try:
                handle = rt.cudaIpcMemHandle_t()
                handle.reserved = handle_bytes[0]
                err, ptr = rt.cudaIpcOpenMemHandle(handle, rt.cudaIpcMemLazyEnablePeerAccess)
                self._check(err, "cudaIpcOpenMemHandle")
                self.ptr = int(ptr)
            except Exception as exc:
                ok, error = 0, exc
        # agree on the outcome: a rank that could not map the buffer makes EVERY rank raise (nobody is left
        # waiting in a later barrier), after the ranks that did succeed have released their mapping
        flags = [None] * self.world
        dist.all_gather_object(flags, ok)
        if not all(flags):
            self._views = 0
            self.close(barrier=False)
            raise RuntimeError("PeerGatherBuffer: rank(s) %s could not map the buffer of rank %d%s"
                               % ([r for r, f in enumerate(flags) if not f], dst, ": %s" % error if error else ""))
        self._views = 0

    @staticmethod
    def _check(err, what):
        if int(err) != 0:
            raise RuntimeError("%s failed: %s" % (what, err))

    def _wrap(self, offset_bytes, shape):
        torch = self._torch
        typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4"}[self.dtype]

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (self.ptr + offset_bytes, False),
                                        "version": 3, "strides": None}
        t = torch.as_tensor(raw, device="cuda")
        t._peer_buffer_keepalive = self
        self._views += 1          # close() refuses to unmap while tensors over the mapping may be alive
        import weakref
        weakref.finalize(t, self._release_view)
        return t

    def _release_view(self):
        self._views -= 1

    def view(self, rank=None):
        """The block of `rank` (default: this rank) inside the peer buffer, as a CUDA tensor."""
        r = self.rank if rank is None else rank
        block = self.rows * self.row_elems * self.itemsize
        return self._wrap(r * block, (self.rows,) + self.row_shape)

    def push(self, local):
        """Copy this rank's finished block into its slot of the peer buffer with one device-to-peer copy
        (large NVLink writes).  Measured faster than letting the kernel store remotely and than an NCCL
        gather: 8 GPUs 1.40 ms against 2.70 ms and 1.69 ms (profiles/r1_peer_gather.txt)."""
        self.view().copy_(local)

    def whole(self):
        """All blocks, rank-major (meaningful on the owner after finish())."""
        return self._wrap(0, (self.world * self.rows,) + self.row_shape)

    def finish(self):
        self._torch.cuda.synchronize()
        self._dist.barrier()

    def close(self, barrier=True):
        rt = self._rt
        self._torch.cuda.synchronize()
        if barrier:
            self._dist.barrier()
        if not self.ptr:
            return
        if getattr(self, "_views", 0) > 0:
            import gc
            gc.collect()
            if self._views > 0:
                raise RuntimeError("PeerGatherBuffer.close(): %d tensor(s) returned by view() / whole() are still alive; "
                                   "drop them first (they alias the mapping that close() releases)" % self._views)
        if self._owner:
            rt.cudaFree(self.ptr)
        else:
            rt.cudaIpcCloseMemHandle(self.ptr)
        self.ptr = 0
