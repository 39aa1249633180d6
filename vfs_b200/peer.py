"""Python handle of the peer-memory communicator (csrc/comm.cu, ``vfs_comm_*`` in include/vfs_b200.h).

torch.distributed is only used ONCE, to move the CUDA IPC handles of the symmetric segments between the ranks
(plumbing); the SyncBN statistic exchanges, the logged-scalar reduction and the gradient all-reduce of the training
step are kernels of libvfs_b200.so over NVLink peer memory, enqueued on the current stream -- no NCCL call, no host
round trip, CUDA-graph capturable.  ``install()`` makes a communicator the process-wide one that ``ops.bn_finalize`` /
``ops.bn_backward`` / ``BaseTracker._parse_losses`` / ``dp.FlatGrads`` pick up; without one those fall back to
``torch.distributed`` collectives (the gloo CPU tests and single-rank runs).
"""
import ctypes

import torch

from . import _native as nat
from ._native import current_stream, ptr

_ACTIVE = None


class _DeviceBlob:
    """Minimal ``__cuda_array_interface__`` owner so torch can view library-owned device memory without copying."""

    def __init__(self, address, nbytes, keepalive):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes), ), typestr='|u1', data=(int(address), False),
                                             version=2, strides=None)
        self._keepalive = keepalive


class PeerComm:

    def __init__(self, data_bytes=0, device=None, rank=None, world=None, exchange=None):
        """``data_bytes``: size of the symmetric data region (the flat gradient buffer lives there).  ``exchange``:
        callable(list-of-bytes-for-this-rank) -> list over ranks; defaults to torch.distributed.all_gather_object."""
        import torch.distributed as dist
        if rank is None:
            rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if not torch.cuda.is_available():
            raise RuntimeError('vfs_b200.peer.PeerComm needs a CUDA device (no CPU fallback)')
        self.rank, self.world = int(rank), int(world)
        self.device = torch.device('cuda', torch.cuda.current_device()) if device is None else torch.device(device)
        lib = nat.lib()
        hbytes = lib.vfs_comm_handle_bytes()
        handle = ctypes.create_string_buffer(hbytes)
        self._c = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(lib.vfs_comm_create(self.rank, self.world, int(data_bytes), ctypes.byref(self._c), handle),
                      'comm_create')
            if self.world > 1:
                mine = bytes(handle.raw)
                if exchange is None:
                    gathered = [None] * self.world
                    dist.all_gather_object(gathered, mine)
                else:
                    gathered = exchange(mine)
                blob = b''.join(gathered)
                assert len(blob) == hbytes * self.world
                buf = ctypes.create_string_buffer(blob, len(blob))
                nat.check(lib.vfs_comm_connect(self._c, ctypes.cast(buf, ctypes.c_void_p)), 'comm_connect')
                dist.barrier()
        self.data_bytes = int(lib.vfs_comm_data_bytes(self._c))
        self._data_ptr = int(lib.vfs_comm_data_ptr(self._c) or 0)
        self._data = None

    # ------------------------------------------------------------------ memory
    def data(self):
        """uint8 view of the whole symmetric data region (same offsets on every rank)."""
        if self._data is None:
            if self.data_bytes == 0:
                raise RuntimeError('PeerComm was created without a data region')
            self._data = torch.as_tensor(_DeviceBlob(self._data_ptr, self.data_bytes, self), device=self.device)
        return self._data

    def offset_of(self, t):
        off = t.data_ptr() - self._data_ptr
        if off < 0 or off + t.numel() * t.element_size() > self.data_bytes:
            raise RuntimeError('tensor does not live in the symmetric data region')
        return off

    # ------------------------------------------------------------------ collectives (enqueued on the current stream)
    def allreduce_small_(self, t):
        """In-place sum over ranks of a small contiguous fp64 / fp32 CUDA tensor (<= 32 KB)."""
        assert t.is_cuda and t.is_contiguous()
        lib = nat.lib()
        if t.dtype == torch.float64:
            rc = lib.vfs_comm_allreduce_small_f64(self._c, ptr(t), t.numel(), current_stream())
        elif t.dtype == torch.float32:
            rc = lib.vfs_comm_allreduce_small_f32(self._c, ptr(t), t.numel(), current_stream())
        else:
            raise TypeError(f'allreduce_small_: dtype {t.dtype} unsupported')
        from . import ops
        ops.check(rc, 'comm_allreduce_small')
        return t

    def allreduce_(self, t, scale=1.0):
        """In-place sum (times ``scale``) over ranks of a float32 tensor living in the data region."""
        assert t.dtype == torch.float32 and t.is_contiguous()
        from . import ops
        ops.check(nat.lib().vfs_comm_allreduce_f32(self._c, self.offset_of(t), t.numel(), float(scale),
                                                   current_stream()), 'comm_allreduce_f32')
        return t

    def barrier(self):
        from . import ops
        ops.check(nat.lib().vfs_comm_barrier(self._c, current_stream()), 'comm_barrier')

    def check(self):
        """Raise if a device-side wait timed out (synchronises)."""
        torch.cuda.synchronize(self.device)
        if nat.lib().vfs_comm_error(self._c):
            raise RuntimeError('vfs_b200 peer communicator: a device-side wait timed out (a rank died or the ranks '
                               'issued different collective sequences)')

    def close(self):
        global _ACTIVE
        if self._c:
            self._data = None
            nat.lib().vfs_comm_destroy(self._c)
            self._c = ctypes.c_void_p()
        if _ACTIVE is self:
            _ACTIVE = None


def install(comm):
    global _ACTIVE
    _ACTIVE = comm
    return comm


def active():
    """The installed communicator, or None."""
    return _ACTIVE
