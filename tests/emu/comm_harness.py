"""Drive the peer-memory collectives of ``csrc/fedcomm.cu`` under the SIMT emulator: R "ranks" in one process.

Every rank is a stream key of the emulator's deferred mode; its kernels are queued by calling the library's own
``extern "C"`` entry points (``flpr_comm_*``) with host buffers, then ``run()`` executes all ranks concurrently under a
seeded random schedule (``cuda_emu.h``). "Host work" between collectives (a client writing its next upload, somebody
archiving a result) is expressed as kernels too - ``flpr_comm_pull_copy`` at world 1 on a private flag page - because in
deferred mode only queued kernels are ordered against each other, exactly like device work on a stream."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import torch

P, I, D, Z, F = C.c_void_p, C.c_int, C.c_double, C.c_size_t, C.c_float


def load(path: str) -> C.CDLL:
    lib = C.CDLL(path)
    sig = {
        "flpr_comm_barrier": [I, I, P, D, P],
        "flpr_comm_reduce_bcast": [I, I, P, D, I, P, P, P, P, Z, I, P],
        "flpr_comm_reduce_bcast_nvls": [I, I, P, D, I, P, P, P, I, P, F, P, P, P, Z, I, P],
        "flpr_comm_mix": [I, I, P, D, I, I, P, P, P, P, P, P, Z, I, P],
        "flpr_comm_curv_moments": [I, I, P, D, I, P, P, P, P, P, Z, I, P],
        "flpr_comm_gather_strided": [I, I, P, D, I, P, P, Z, I, P],
        "flpr_comm_pull_copy": [I, I, P, D, P, P, P, Z, I, P],
        "flpr_comm_set_channel": [I],
        "flpr_emu_run": [C.c_uint, C.c_long, C.c_uint],
        "flpr_emu_set_start_delay": [P, I],
        "flpr_emu_set_lane_slowdown": [P, C.c_uint],
        "flpr_emu_mc_register": [P, Z, I, P],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = I
    lib.flpr_comm_set_one_shot_bytes.argtypes = [I]
    lib.flpr_comm_set_one_shot_bytes.restype = None
    lib.flpr_emu_defer.argtypes = [I]
    lib.flpr_emu_defer.restype = None
    lib.flpr_emu_mc_clear.argtypes = []
    lib.flpr_emu_mc_clear.restype = None
    lib.flpr_emu_clock_ns.restype = C.c_ulonglong
    for name in ("flpr_comm_flag_page_bytes", "flpr_emu_deadlocks", "flpr_comm_max_channels"):
        getattr(lib, name).restype = I
        getattr(lib, name).argtypes = []
    return lib


def ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    return (P * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])


class EmuWorld:
    """``world`` ranks sharing one address space; ``pages[r]`` is rank r's flag page (what the symmetric arena starts
    with), ``priv[r]`` a private page for the world-1 helper kernels of rank r."""

    def __init__(self, lib: C.CDLL, world: int, blocks: int = 2, timeout_s: float = 1e3):
        self.lib, self.world, self.blocks, self.timeout_s = lib, world, blocks, timeout_s
        words = lib.flpr_comm_flag_page_bytes() // 4
        self.err_off = words - 4
        self.pages = [torch.zeros(words, dtype=torch.int32) for _ in range(world)]
        self.priv = [torch.zeros(words, dtype=torch.int32) for _ in range(world)]
        self._pages = ptr_array(self.pages)
        self._keep: List[object] = []          # lives until close()
        self._keep_run: List[object] = []      # lives until the next run()
        lib.flpr_emu_defer(1)

    # -- addressing ------------------------------------------------------------------------------------------------------
    def stream(self, rank: int, idx: int = 0) -> P:
        return P(0x100000 + rank * 0x100 + idx * 0x10)       # any distinct non-null handle: the emulator's queue key

    def error_word(self, rank: int) -> int:
        return int(self.pages[rank][self.err_off])

    def set_start_delay(self, rank: int, passes: int, idx: int = 0) -> None:
        self.lib.flpr_emu_set_start_delay(self.stream(rank, idx), passes)

    def set_slowdown(self, rank: int, one_in: int, idx: int = 0) -> None:
        """Rank ``rank``'s threads run in one of ``one_in`` scheduler passes only (a slow GPU / a busy SM)."""
        self.lib.flpr_emu_set_lane_slowdown(self.stream(rank, idx), one_in)

    def _check(self, rc: int, what: str) -> None:
        assert rc == 0, f"{what} returned {rc}"

    # -- collectives (one call = this rank's launch) ---------------------------------------------------------------------
    def reduce_bcast(self, rank, src, dst, cnt=None, w=None, stream=0, channel=0, blocks=None):
        n = src[0].numel()
        wv = None if w is None else (F * len(w))(*w)
        self.lib.flpr_comm_set_channel(channel)
        self._check(self.lib.flpr_comm_reduce_bcast(
            rank, self.world, self._pages, self.timeout_s, len(src), ptr_array(src),
            None if cnt is None else ptr_array(cnt), wv, ptr_array(dst), n, blocks or self.blocks,
            self.stream(rank, stream)), "reduce_bcast")

    def reduce_bcast_nvls(self, rank, local_src, local_cnt, local_w, cnt_all, w_total, partial, mc_partial, mc_dst,
                          n_clients, stream=0, channel=0, blocks=None):
        n = partial.numel()
        wv = None if local_w is None else (F * max(len(local_w), 1))(*local_w)
        self.lib.flpr_comm_set_channel(channel)
        self._check(self.lib.flpr_comm_reduce_bcast_nvls(
            rank, self.world, self._pages, self.timeout_s, len(local_src), ptr_array(local_src) if local_src else None,
            None if local_cnt is None else ptr_array(local_cnt), wv, n_clients,
            None if cnt_all is None else ptr_array(cnt_all), float(w_total), partial.data_ptr(),
            mc_partial.data_ptr(), mc_dst.data_ptr(), n, blocks or self.blocks, self.stream(rank, stream)),
            "reduce_bcast_nvls")

    def mix(self, rank, src, rows, dst_g, dst_theta, dst_bf16, rows_dev=None, stream=0, channel=0, blocks=None):
        n = src[0].numel()
        L = len(dst_g)
        flat = None
        if rows is not None:
            vals = [float(v) for r in rows for v in r]
            flat = (F * max(len(vals), 1))(*vals)
        self.lib.flpr_comm_set_channel(channel)
        self._check(self.lib.flpr_comm_mix(
            rank, self.world, self._pages, self.timeout_s, len(src), L, ptr_array(src), flat,
            None if rows_dev is None else rows_dev.data_ptr(), ptr_array(dst_g) if L else None,
            ptr_array(dst_theta) if L else None, ptr_array(dst_bf16) if L else None, n, blocks or self.blocks,
            self.stream(rank, stream)), "mix")

    def curv_moments(self, rank, fisher, param, dst_f, dst_fp, dst_fpp, stream=0, channel=0):
        n = fisher[0].numel()
        self.lib.flpr_comm_set_channel(channel)
        self._check(self.lib.flpr_comm_curv_moments(
            rank, self.world, self._pages, self.timeout_s, len(fisher), ptr_array(fisher), ptr_array(param),
            ptr_array(dst_f), ptr_array(dst_fp), ptr_array(dst_fpp), n, self.blocks, self.stream(rank, stream)),
            "curv_moments")

    def gather_strided(self, rank, src, dst, n, stream=0, channel=0):
        self.lib.flpr_comm_set_channel(channel)
        self._check(self.lib.flpr_comm_gather_strided(
            rank, self.world, self._pages, self.timeout_s, len(src), ptr_array(src), dst.data_ptr(), n, self.blocks,
            self.stream(rank, stream)), "gather_strided")

    def pull_copy(self, rank, src, dst=None, dst_bf16=None, stream=0, channel=0):
        self.lib.flpr_comm_set_channel(channel)
        self._check(self.lib.flpr_comm_pull_copy(
            rank, self.world, self._pages, self.timeout_s, src.data_ptr(), None if dst is None else dst.data_ptr(),
            None if dst_bf16 is None else dst_bf16.data_ptr(), src.numel(), self.blocks, self.stream(rank, stream)),
            "pull_copy")

    def barrier(self, rank, stream=0, channel=0):
        self.lib.flpr_comm_set_channel(channel)
        self._check(self.lib.flpr_comm_barrier(rank, self.world, self._pages, self.timeout_s,
                                               self.stream(rank, stream)), "barrier")

    def local_copy(self, rank, src, dst, stream=0):
        """``dst <- src`` as a kernel on rank ``rank``'s stream (world 1, private flags): device-ordered "host work"."""
        one = ptr_array([self.priv[rank]])
        self._keep_run.append(one)
        self.lib.flpr_comm_set_channel(0)
        self._check(self.lib.flpr_comm_pull_copy(0, 1, one, self.timeout_s, src.data_ptr(), dst.data_ptr(), None,
                                                 src.numel(), 1, self.stream(rank, stream)), "local_copy")

    def multicast(self, members: Sequence[torch.Tensor]) -> torch.Tensor:
        """An emulated NVSwitch multicast window over one buffer per rank; returns the tensor whose ADDRESS RANGE plays the
        multicast mapping (its contents are never touched)."""
        window = torch.empty_like(members[0])
        self._keep.append(window)
        self.lib.flpr_emu_mc_register(window.data_ptr(), window.numel() * window.element_size(), len(members),
                                      ptr_array(members))
        return window

    # -- execution -------------------------------------------------------------------------------------------------------
    def run(self, seed: int, max_passes: int = 200000, stall_one_in: int = 4) -> int:
        rc = self.lib.flpr_emu_run(seed, max_passes, stall_one_in)
        self._keep_run.clear()
        return rc

    def close(self) -> None:
        self.lib.flpr_emu_mc_clear()
        self.lib.flpr_emu_defer(0)


# ---------------------------------------------------------------------------------------------------- FedComm on the emulator
def make_fedcomm_world(lib: C.CDLL, world: int, num_clients: int, arena_bytes: int = 8 << 20, blocks: int = 2,
                       multicast: bool = True, timeout_s: float = 1e3):
    """``world`` instances of the product's ``FedComm`` (``parallel/comm.py``) in ONE process, in ``p2p`` mode, on host
    arenas and the emulated kernels: the Python half of the communication layer - buffer placement, owner / slot maps,
    pointer tables, kernel choice (two-shot / one-shot / NVLS), grid sizes, the launch loop of the mix - runs unchanged;
    only the arena construction (VMM / IPC) and the stream handle are replaced. Every collective must be called on all
    ranks, then ``run(lib, seed)`` executes what was queued."""
    from flpr_b200.ops import native
    from flpr_b200.parallel.comm import FedComm

    native.declare_present(lib)
    lib.flpr_emu_defer(1)
    flag_bytes = ((lib.flpr_comm_flag_page_bytes() + 4095) // 4096) * 4096
    arenas = [torch.zeros(arena_bytes, dtype=torch.uint8) for _ in range(world)]
    window = torch.empty(arena_bytes, dtype=torch.uint8) if multicast else None
    if multicast:
        lib.flpr_emu_mc_register(window.data_ptr(), arena_bytes, world, ptr_array(arenas))

    class EmuFedComm(FedComm):
        def __init__(self, rank: int):                      # (no torch.distributed, no CUDA: fields set by hand)
            self.device = torch.device("cpu")
            self.group = None
            self.rank, self.world, self.K = rank, world, int(num_clients)
            self.slots = (self.K + world - 1) // world
            self.mode = self.backend = "p2p"
            self.timeout_s = float(timeout_s)
            self.arena_bytes = arena_bytes
            self.bufs = {}
            self._cursor = flag_bytes
            self._keep = []
            self.bytes_moved = 0
            self.nvls, self.nvls_min_bytes, self.nvls_launches = True, 1 << 20, 0
            self.block_cap, self.comm_blocks = 0, blocks
            self._vmm = None
            self._lib = lib
            self._arena = arenas[rank]
            self._base = arenas[rank].data_ptr()
            self._peer_base = [a.data_ptr() for a in arenas]
            self._mc_base = window.data_ptr() if multicast else 0
            self._flag_pages = (C.c_void_p * world)(*[C.c_void_p(b) for b in self._peer_base])
            self._mailbox = torch.zeros(4, dtype=torch.int32)
            self._window = window
            # what flpr_comm_set_mailbox does with a cudaMemcpy: the 64-bit address of the host mailbox at MAILBOX_OFF
            words = lib.flpr_comm_flag_page_bytes() // 4
            page = self._arena[:words * 4].view(torch.int32)
            addr = self._mailbox.data_ptr()
            lo, hi = addr & 0xFFFFFFFF, addr >> 32
            page[words - 2] = lo - (1 << 32) if lo >= (1 << 31) else lo
            page[words - 1] = hi - (1 << 32) if hi >= (1 << 31) else hi

        def _stream(self):
            return P(0x200000 + self.rank * 0x100)

        def error_word(self) -> int:
            words = lib.flpr_comm_flag_page_bytes() // 4
            return int(self._arena[:words * 4].view(torch.int32)[words - 4])

        def close(self) -> None:
            pass

    return [EmuFedComm(r) for r in range(world)]


def run(lib: C.CDLL, seed: int, max_passes: int = 400000, stall_one_in: int = 4) -> int:
    return lib.flpr_emu_run(seed, max_passes, stall_one_in)


def end(lib: C.CDLL) -> None:
    lib.flpr_emu_mc_clear()
    lib.flpr_emu_defer(0)
