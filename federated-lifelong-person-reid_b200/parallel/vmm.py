"""Symmetric arena on the CUDA virtual-memory-management API, with an NVSwitch multicast (NVLS) view of it.

``cudaMalloc`` + CUDA IPC (the first arena back-end of :class:`~.comm.FedComm`) gives peer *unicast* addresses only. The
NVSwitch can do more: a **multicast object** bound to one physical allocation per GPU yields an address on which
``multimem.st`` writes to every GPU and ``multimem.ld_reduce`` returns the element-wise sum over all GPUs, computed in
the switch. Binding needs allocations made with ``cuMemCreate``, so this module builds the whole symmetric arena on
the VMM API:

* every rank ``cuMemCreate``-s its arena (device-pinned, exportable as a POSIX file descriptor), maps it locally and
  exports the handle; the file descriptors travel between the rank processes over ``AF_UNIX`` sockets
  (``SCM_RIGHTS``); every rank imports and maps every peer's handle -> the same ``peer_base[r] + offset`` addressing
  the IPC back-end provides;
* rank 0 creates the multicast object, its descriptor is shared the same way, every rank adds its device, binds its own
  arena at offset 0 and maps the object -> ``mc_base + offset`` is the multicast address of the symmetric buffer at
  ``offset``.

Everything goes through ``cuda.bindings.driver`` (cuda-python); no process ever needs another's virtual addresses.
Any failure (no multicast support, no fabric manager, a driver error) raises :class:`VmmUnavailable` and the caller
falls back to the IPC arena - the collectives then use their peer-load / peer-store kernels only.
"""
from __future__ import annotations

import os
import socket
import struct
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


class VmmUnavailable(RuntimeError):
    pass


def _drv():
    try:
        from cuda.bindings import driver as cu
    except Exception:  # pragma: no cover
        try:
            from cuda import cuda as cu          # older cuda-python layout
        except Exception as ex:
            raise VmmUnavailable(f"cuda-python is not importable: {ex}")
    return cu


def _ck(res, what: str):
    cu = _drv()
    err = res[0]
    if err != cu.CUresult.CUDA_SUCCESS:
        raise VmmUnavailable(f"{what} -> {err}")
    if len(res) == 1:
        return None
    return res[1] if len(res) == 2 else res[1:]


def _exchange_fds(mine: List[int], rank: int, world: int, group, tag: str) -> List[List[int]]:
    """All-gather of file descriptors between the rank processes of one node: ``out[r]`` = rank r's descriptors
    (duplicated into this process). ``AF_UNIX`` abstract-namespace sockets + ``SCM_RIGHTS``."""
    token = [f"{os.getpid()}-{int.from_bytes(os.urandom(4), 'little')}"]
    dist.broadcast_object_list(token, src=0, group=group)
    addr = lambda r: f"\0flpr-{token[0]}-{tag}-{r}"            # noqa: E731  (abstract namespace: no file to clean up)
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(addr(rank))
    srv.listen(64)
    dist.barrier(group=group)
    conns = []
    try:
        for p in range(world):
            if p == rank:
                continue
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(addr(p))
            socket.send_fds(c, [struct.pack("<i", rank)], list(mine))
            conns.append(c)
        out: List[Optional[List[int]]] = [None] * world
        out[rank] = list(mine)
        srv.settimeout(120.0)
        for _ in range(world - 1):
            conn, _ = srv.accept()
            msg, fds, _, _ = socket.recv_fds(conn, 16, len(mine))
            if len(fds) != len(mine):
                raise VmmUnavailable(f"expected {len(mine)} descriptors, received {len(fds)}")
            out[struct.unpack("<i", msg[:4])[0]] = list(fds)
            conn.close()
        dist.barrier(group=group)                            # every send has been received before sockets close
    finally:
        for c in conns:
            c.close()
        srv.close()
    return out  # type: ignore[return-value]


class SymmetricVmm:
    """``base`` (local arena), ``peer_base[r]`` (rank r's arena mapped here) and ``mc_base`` (multicast view, 0 when
    the switch cannot multicast)."""

    def __init__(self, device: torch.device, nbytes: int, rank: int, world: int, group=None, multicast: bool = True):
        cu = _drv()
        self.cu = cu
        self.rank, self.world = rank, world
        dev = device.index if device.index is not None else torch.cuda.current_device()
        torch.zeros(1, device=device)                              # the primary context exists and is current
        _ck(cu.cuInit(0), "cuInit")
        cudev = _ck(cu.cuDeviceGet(dev), "cuDeviceGet")
        A = cu.CUdevice_attribute
        if not _ck(cu.cuDeviceGetAttribute(A.CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, cudev),
                   "query posix-fd handles"):
            raise VmmUnavailable("POSIX file-descriptor memory handles are not supported on this device")
        want_mc = multicast and world > 1 and bool(_ck(cu.cuDeviceGetAttribute(
            A.CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, cudev), "query multicast"))
        HT = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
        prop = cu.CUmemAllocationProp()
        prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = dev
        prop.requestedHandleTypes = HT
        gran = int(_ck(cu.cuMemGetAllocationGranularity(
            prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "allocation granularity"))
        mcprop = None
        if want_mc:
            mcprop = cu.CUmulticastObjectProp()
            mcprop.numDevices = world
            mcprop.handleTypes = HT
            mcprop.flags = 0
            mcprop.size = ((nbytes + gran - 1) // gran) * gran
            try:
                mgran = int(_ck(cu.cuMulticastGetGranularity(
                    mcprop, cu.CUmulticastGranularity_flags.CU_MULTICAST_GRANULARITY_RECOMMENDED),
                    "multicast granularity"))
                gran = max(gran, mgran)
            except VmmUnavailable:
                want_mc = False
        # every rank must take the same decision: multicast only if every device supports it
        flag = torch.tensor([1 if want_mc else 0], device=device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        want_mc = bool(flag.item())
        self.size = ((nbytes + gran - 1) // gran) * gran
        self.gran = gran
        self.device = device
        self.group = group
        self._maps: List[int] = []
        self._handles: list = []
        self.peer_base: List[int] = [0] * world
        self.mc_base = 0
        self.mc_handle = None
        self.mc_error = ""

        # Every phase that can fail on one rank only ends in an agreement (all-reduce of a status flag) instead of a
        # bare barrier: a rank that raised would otherwise leave its peers waiting in the next collective for ever.
        def phase(what, fn):
            err = None
            try:
                fn()
            except Exception as ex:  # noqa: BLE001
                err = f"{what}: {type(ex).__name__}: {ex}"
            if not self._agree(err is None):
                raise VmmUnavailable(err or f"{what}: failed on a peer rank")

        def local_alloc():
            self.handle = _ck(cu.cuMemCreate(self.size, prop, 0), "cuMemCreate")
            self._handles.append(self.handle)
            self.access = cu.CUmemAccessDesc()
            self.access.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
            self.access.location.id = dev
            self.access.flags = cu.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
            self.base = self._map(self.handle)
            self.peer_base[rank] = self.base
        phase("local arena", local_alloc)
        if world > 1:
            box = {}

            def export():
                box["fd"] = int(_ck(cu.cuMemExportToShareableHandle(self.handle, HT, 0), "cuMemExportToShareableHandle"))
            phase("export", export)
            box["all"] = _exchange_fds([box["fd"]], rank, world, group, "mem")

            def import_peers():
                for r in range(world):
                    if r == rank:
                        continue
                    h = _ck(cu.cuMemImportFromShareableHandle(box["all"][r][0], HT), f"import arena of rank {r}")
                    self._handles.append(h)
                    self.peer_base[r] = self._map(h)
            try:
                phase("peer arenas", import_peers)
            finally:
                for lst in box["all"]:
                    for f in lst:
                        try:
                            os.close(f)
                        except OSError:
                            pass
        if want_mc:
            try:
                self._setup_multicast(mcprop, cudev, HT, group, phase)
            except VmmUnavailable as ex:                          # agreed on by every rank: unicast arena only
                self.mc_error = str(ex)
                self.mc_base = 0

    # ------------------------------------------------------------------ helpers
    def _agree(self, ok: bool) -> bool:
        if self.world == 1:
            return ok
        t = torch.tensor([1 if ok else 0], device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=self.group)
        return bool(t.item())

    def _map(self, handle) -> int:
        cu = self.cu
        ptr = _ck(cu.cuMemAddressReserve(self.size, self.gran, 0, 0), "cuMemAddressReserve")
        _ck(cu.cuMemMap(ptr, self.size, 0, handle, 0), "cuMemMap")
        _ck(cu.cuMemSetAccess(ptr, self.size, [self.access], 1), "cuMemSetAccess")
        self._maps.append(int(ptr))
        return int(ptr)

    def _setup_multicast(self, mcprop, cudev, HT, group, phase) -> None:
        cu = self.cu
        mcprop.size = self.size
        box = {"fds": [-1]}

        def create():
            if self.rank == 0:
                self.mc_handle = _ck(cu.cuMulticastCreate(mcprop), "cuMulticastCreate")
                box["fds"] = [int(_ck(cu.cuMemExportToShareableHandle(self.mc_handle, HT, 0),
                                      "export multicast object"))]
            else:
                box["fds"] = [os.open(os.devnull, os.O_RDONLY)]          # placeholder: the exchange is an all-gather
        phase("multicast object", create)
        got = _exchange_fds(box["fds"], self.rank, self.world, group, "mc")

        def join():
            if self.rank != 0:
                self.mc_handle = _ck(cu.cuMemImportFromShareableHandle(got[0][0], HT), "import multicast object")
            _ck(cu.cuMulticastAddDevice(self.mc_handle, cudev), "cuMulticastAddDevice")
        try:
            phase("multicast join", join)                     # (agreement = every device is part of the object)
        finally:
            for lst in got:
                for f in lst:
                    try:
                        os.close(f)
                    except OSError:
                        pass
        phase("multicast bind", lambda: _ck(cu.cuMulticastBindMem(self.mc_handle, 0, self.handle, 0, self.size, 0),
                                            "cuMulticastBindMem"))
        box2 = {}
        phase("multicast map", lambda: box2.__setitem__("p", self._map(self.mc_handle)))
        self.mc_base = box2["p"]

    def close(self) -> None:
        cu = self.cu
        try:
            if self.mc_handle is not None and self.mc_base:
                cu.cuMulticastUnbind(self.mc_handle, self.cu.cuDeviceGet(self.access.location.id)[1], 0, self.size)
        except Exception:
            pass
        for p in self._maps:
            try:
                cu.cuMemUnmap(p, self.size)
                cu.cuMemAddressFree(p, self.size)
            except Exception:
                pass
        self._maps = []
        for h in self._handles + ([self.mc_handle] if self.mc_handle is not None else []):
            try:
                cu.cuMemRelease(h)
            except Exception:
                pass
        self._handles = []
