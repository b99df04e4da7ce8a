"""Native NCCL communicator for ``prv2_reduce_canvas`` (the C-ABI form of the sharded path's one exchange).

``torch.distributed`` does not hand out its ``ncclComm_t``, so a caller that wants the reduce issued by the library itself
(a C++ host would) creates one here: rank 0 draws an ``ncclUniqueId``, the existing process group broadcasts its 128 bytes, every
rank calls ``ncclCommInitRank``.  The symbols come from the libnccl PyTorch has already loaded.  Opt-in
(``PatchRefiner.use_native_reduce()`` or ``PRV2_NATIVE_REDUCE=1``); the default remains ``torch.distributed.all_reduce`` on the
process group the host framework owns -- both are one NCCL sum all-reduce of the same buffer."""
from __future__ import annotations

import ctypes as C

import torch


class _UniqueId(C.Structure):
    _fields_ = [("internal", C.c_char * 128)]


class NcclComm:
    def __init__(self, device: torch.device):
        dist = torch.distributed
        assert dist.is_initialized(), "create the process group first (the unique id travels over it)"
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._nccl = C.CDLL("libnccl.so.2")
        uid = _UniqueId()
        if self.rank == 0:
            self._check(self._nccl.ncclGetUniqueId(C.byref(uid)), "ncclGetUniqueId")
        on_dev = dist.get_backend() == "nccl"
        t = torch.frombuffer(bytearray(bytes(uid)), dtype=torch.uint8).clone()
        t = t.to(device) if on_dev else t
        dist.broadcast(t, src=0)
        C.memmove(C.byref(uid), bytes(t.cpu().numpy().tobytes()), 128)
        self.handle = C.c_void_p()
        self._nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _UniqueId, C.c_int]
        with torch.cuda.device(device):
            self._check(self._nccl.ncclCommInitRank(C.byref(self.handle), self.world, uid, self.rank), "ncclCommInitRank")

    def _check(self, rc: int, what: str):
        if rc != 0:
            self._nccl.ncclGetErrorString.restype = C.c_char_p
            raise RuntimeError(f"{what} failed: {self._nccl.ncclGetErrorString(rc).decode()}")

    def destroy(self):
        if self.handle:
            self._nccl.ncclCommDestroy.argtypes = [C.c_void_p]
            self._nccl.ncclCommDestroy(self.handle)
            self.handle = C.c_void_p()
