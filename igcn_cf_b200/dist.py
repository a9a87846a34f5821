"""Multi-GPU plumbing: one process per GPU on one NVSwitch box.

The reference is single-process / single-GPU, so nothing here mirrors a reference
interface.  What shards, and how:

* propagation -- every rank owns one slice of the user rows and one of the item rows, each balanced by
  non-zeros (`shard_bounds`, applied per half by graph._row_ranges); every rank keeps a full replica of
  each layer's input and computes its own rows; the kernel epilogue stores the finished rows into EVERY
  rank's copy through peer-mapped memory (the all-gather is fused into the SpMM; large blocks go out
  through igcn_peer_push instead), and `PeerContext.barrier()` -- a device-side flag barrier on the
  launch stream -- separates the layers.  The result is bit-identical to one GPU: a row is reduced by
  one rank in the same order whatever the rank count.  Whether rows are sharded at all is the model's
  decision (model_config['shard'], 'auto' = only when the exchange pays off).
* BPR step / Adam -- replicated on the gathered representation (6,144 rows; cheaper than talking).
* evaluation -- users split evenly (`split_range`), no communication until the top-k lists are
  gathered for the metrics (`gather_rows`).

torch.distributed is only used for rendezvous (exchanging IPC handles) and the final list gather.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import call, stream_ptr

_CONTEXT = None


def shard_bounds(rowptr, world, row_cost=4):
    """Row offsets [world + 1] of contiguous blocks with about equal nnz + row_cost * rows."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    n = len(rowptr) - 1
    cost = rowptr + row_cost * np.arange(n + 1, dtype=np.int64)
    targets = cost[-1] * np.arange(1, world, dtype=np.float64) / world
    cuts = np.searchsorted(cost, targets, side='left')
    bounds = np.concatenate([[0], cuts, [n]]).astype(np.int64)
    return np.maximum.accumulate(bounds)


def split_range(n, rank, world):
    """[lo, hi) of `rank` when n items are dealt out in `world` contiguous, near-equal pieces."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, n_total, group=None):
    """Concatenate per-rank row blocks (split_range order) of a 2-D tensor on every rank."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    width = (int(n_total) + world - 1) // world
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, padded, group=group)
    parts = []
    for r in range(world):
        lo, hi = split_range(n_total, r, world)
        parts.append(out[r * width:r * width + (hi - lo)])
    return torch.cat(parts, dim=0)


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view a cudaMalloc'ed region."""

    def __init__(self, ptr_, shape, typestr):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr_), False),
                                         'version': 2, 'strides': None}


_TYPESTR = {torch.float32: '<f4', torch.int32: '<i4', torch.uint8: '|u1', torch.int64: '<i8'}


class PeerBuf:
    """One symmetric buffer: `tensor` is the local copy, `ptrs[r]` its base address on rank r as
    mapped into THIS process (ptrs[rank] == tensor.data_ptr())."""

    def __init__(self, tensor, ptrs):
        self.tensor, self.ptrs = tensor, list(ptrs)
        self._arrays = {}

    def peer_array(self, byte_offset=0):
        arr = self._arrays.get(byte_offset)
        if arr is None:
            arr = (C.c_void_p * len(self.ptrs))(*[p + byte_offset for p in self.ptrs])
            self._arrays[byte_offset] = arr
        return arr


class PeerContext:
    """Rank/world of this process plus the symmetric-buffer allocator and the device barrier."""

    def __init__(self, group=None, device=None):
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError('igcn_cf_b200.dist: torch.distributed is not initialised')
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > _lib.MAX_PEERS:
            raise RuntimeError('at most %d ranks (one NVSwitch box) are supported' % _lib.MAX_PEERS)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        self._owned, self._mapped = [], []
        self.flags = self.alloc((_lib.MAX_PEERS,), torch.int32)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.n_barriers = 0

    def alloc(self, shape, dtype=torch.float32):
        """Collective: every rank must call it with the same shape in the same order."""
        import torch.distributed as dist
        shape = tuple(int(s) for s in shape)
        nbytes = max(256, int(np.prod(shape)) * torch.empty(0, dtype=dtype).element_size())
        p = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        with torch.cuda.device(self.device):
            call('igcn_peer_alloc', nbytes, C.byref(p), C.cast(handle, C.c_void_p))
            self._owned.append(p.value)
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=self.group)
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(p.value)
                    continue
                q = C.c_void_p()
                buf = (C.c_uint8 * 64).from_buffer_copy(h)
                call('igcn_peer_open', C.cast(buf, C.c_void_p), C.byref(q))
                self._mapped.append(q.value)
                ptrs.append(q.value)
            t = torch.as_tensor(_RawCuda(p.value, shape, _TYPESTR[dtype]), device=self.device)
        return PeerBuf(t, ptrs)

    def barrier(self):
        """Device-side barrier on the current stream (capturable); no host synchronisation."""
        call('igcn_peer_barrier', self.flags.peer_array(), self.world, self.rank, self.epoch.data_ptr(),
             self.status.data_ptr(), stream_ptr())
        self.n_barriers += 1

    def check(self):
        """Host-side: raise if any barrier so far timed out (one D2H read)."""
        if int(self.status.item()) != 0:
            raise RuntimeError('igcn_cf_b200.dist: a peer did not reach a barrier within the timeout')

    def close(self):
        torch.cuda.synchronize(self.device)
        for q in self._mapped:
            call('igcn_peer_close', q)
        for p in self._owned:
            call('igcn_peer_free', p)
        self._mapped, self._owned = [], []


def init_peers(group=None, device=None):
    """Create (once) the process-wide PeerContext; returns None when the world has one rank."""
    global _CONTEXT
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return None
    if _CONTEXT is None:
        _CONTEXT = PeerContext(group, device)
    return _CONTEXT


def current():
    return _CONTEXT


def shutdown():
    global _CONTEXT
    if _CONTEXT is not None:
        _CONTEXT.close()
        _CONTEXT = None
