"""Fused Ulysses exchange over NVLink peer memory (the B200-native replacement of the four NCCL
all-to-alls per layer in models/wan/distributed/ulysses.py:32-46).

Every rank owns one exchange buffer (allocated and IPC-exported through the C ABI, include/univid_b200.h
`uvb_sp_*`) laid out as

    flags (4 KiB) | q_recv | k_recv | v_recv  each [B, p, s, n, 128]  ==  [B, L, n, 128]
                  | o_recv                         [B, s, N, 128]

with p ranks, s = L/p local tokens, n = N/p local heads.  Rank i's producer kernels (q/k RMSNorm+RoPE
prologue, v head scatter) store head group j directly into slot (b, i) of rank j's q/k/v_recv; rank j's
attention kernel reads [B, L, n, 128] in place and TMA-stores each output tile into o_recv of the rank that
owns its tokens, at head offset j*n.  Two flag words per peer pair order producers and consumers
(`uvb_sp_signal` / `uvb_sp_wait`); the value is the exchange counter, identical on all ranks because every
rank executes the same sequence of layers.  No NCCL call, no pack/unpack pass and no intermediate copy is
left on the data path; torch.distributed is used once, to swap the 64-byte IPC handles.

Buffer reuse is safe without double buffering: a rank only starts the producers of exchange e+1 after it
has waited for every peer's o-signal of exchange e (so all attention kernels that read q/k/v_recv are done),
and peers only store o of exchange e+1 after this rank's qkv-signal e+1, which it issues after consuming
o_recv of exchange e (same stream).  `attend()` therefore returns a VIEW of o_recv that must be consumed on
the current stream before the next exchange (the o projection does exactly that).
"""
import atexit
import ctypes
import os
import warnings

import torch
import torch.distributed as dist

from ... import _ext

_FLAG_BYTES = 4096
_O_FLAG_OFF = 128          # bytes: qkv flags at [0, 4p), o flags at [128, 128 + 4p)
_CONTEXTS = {}
_DISABLED = None


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can alias memory the C ABI allocated."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False),
                                         "version": 2}


def _align(n, a=256):
    return (n + a - 1) // a * a


def exchange_layout(B, s, N, p, rank):
    """Pure layout arithmetic of one rank's view of the exchange (unit-tested on the CPU against the oracle's
    emulation of the reference all_to_all, tests/test_p2p_layout_cpu.py).  Offsets into an exchange buffer are
    bytes; strides are bf16 elements.
      q/k/v_recv of rank j = [B, p, s, n, 128]: rank `rank` stores element (b, l, h, d) of its head group j at
          off_x + slot_bytes + 2 * (b * send_sb + l * send_sl + h * 128 + d)        inside rank j's buffer
      o_recv of rank i = [B, s, N, 128]: rank `rank` stores rows [i*s, (i+1)*s) of its output at head offset
          rank * n
      flags: this rank's word inside every peer's flag area (qkv at byte 4*rank, o at 128 + 4*rank)."""
    n = N // p
    elems = B * s * N * 128                       # == B * L * n * 128
    off_q = _FLAG_BYTES
    off_k = off_q + _align(elems * 2)
    off_v = off_k + _align(elems * 2)
    off_o = off_v + _align(elems * 2)
    return dict(n=n, elems=elems, off_q=off_q, off_k=off_k, off_v=off_v, off_o=off_o,
                nbytes=off_o + _align(elems * 2), slot_bytes=rank * s * n * 128 * 2,
                send_sb=p * s * n * 128, send_sl=n * 128, o_head_offset=rank * n,
                qkv_flag_bytes=4 * rank, o_flag_bytes=_O_FLAG_OFF + 4 * rank)


class PeerExchangeUnavailable(RuntimeError):
    """Raised on EVERY rank when any rank could not allocate / export / map the exchange buffers."""


class UlyssesP2P:
    def __init__(self, B, s, N, device, group=None):
        """Collective: every rank of `group` must call it.  Set-up runs in two phases (allocate + export, import) and
        the ranks agree on success after each with an all_reduce(MIN), so that a failure on ONE rank (cudaMalloc,
        cudaIpcGetMemHandle, cudaIpcOpenMemHandle for one peer) makes EVERY rank raise PeerExchangeUnavailable at the
        same point, after executing the same collectives -- nobody is left waiting in a barrier -- and whatever was
        allocated or mapped so far is released."""
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        p = self.world
        if N % p != 0:
            raise ValueError(f"{N} heads cannot be split over {p} ranks")
        self.B, self.s, self.N, self.n, self.L = B, s, N, N // p, p * s
        self.device = device
        lay = exchange_layout(B, s, N, p, self.rank)
        elems = lay["elems"]
        self.off_q, self.off_k, self.off_v, self.off_o = lay["off_q"], lay["off_k"], lay["off_v"], lay["off_o"]
        self.nbytes = lay["nbytes"]
        self.base, self.peer_base, self._imported, self._closed = 0, [], [], False
        lib = _ext.lib()

        def agree(ok, what, err):
            flag = torch.tensor([1.0 if ok else 0.0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if flag.item() == 0:
                self.close()
                raise PeerExchangeUnavailable(f"{what} failed on " + ("this rank: " + str(err) if not ok else "a peer rank"))

        with torch.cuda.device(device):
            # ---- phase 1: allocate + export
            err, handle = None, ctypes.create_string_buffer(64)
            try:
                base = ctypes.c_void_p()
                _ext._check(lib.uvb_sp_buffer_alloc(self.nbytes, ctypes.byref(base)))
                self.base = int(base.value)
                _ext._check(lib.uvb_sp_ipc_export(self.base, handle))
            except RuntimeError as e:
                err = e
            agree(err is None, "exchange buffer allocation / IPC export", err)
            handles = [None] * p
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            # ---- phase 2: map every peer's buffer
            err = None
            try:
                for j in range(p):
                    if j == self.rank:
                        self.peer_base.append(self.base)
                        continue
                    ptr = ctypes.c_void_p()
                    _ext._check(lib.uvb_sp_ipc_import(ctypes.create_string_buffer(handles[j], 64), ctypes.byref(ptr)))
                    self.peer_base.append(int(ptr.value))
                    self._imported.append(int(ptr.value))
            except RuntimeError as e:
                err = e
            agree(err is None, "IPC import of a peer's exchange buffer", err)
        n, r = self.n, self.rank
        slot = lay["slot_bytes"]                      # byte offset of slot (b = 0, i = rank) in a peer's q/k/v_recv
        self.q_peers = _ext.ptr_array([pb + self.off_q + slot for pb in self.peer_base])
        self.k_peers = _ext.ptr_array([pb + self.off_k + slot for pb in self.peer_base])
        self.v_peers = _ext.ptr_array([pb + self.off_v + slot for pb in self.peer_base])
        self.o_peers = _ext.ptr_array([pb + self.off_o for pb in self.peer_base])
        self.qkv_flag_peers = _ext.ptr_array([pb + lay["qkv_flag_bytes"] for pb in self.peer_base])
        self.o_flag_peers = _ext.ptr_array([pb + lay["o_flag_bytes"] for pb in self.peer_base])
        self.send_sb, self.send_sl = lay["send_sb"], lay["send_sl"]      # element strides of a slot inside [B, p, s, n, 128]
        raw = torch.as_tensor(_RawCuda(self.base, self.nbytes), device=device)
        self._raw = raw
        view = lambda off, shape: raw[off:off + elems * 2].view(torch.bfloat16).view(shape)
        self.q_local = view(self.off_q, (B, self.L, n, 128))
        self.k_local = view(self.off_k, (B, self.L, n, 128))
        self.v_local = view(self.off_v, (B, self.L, n, 128))
        self.o_local = view(self.off_o, (B, s, N, 128))
        self.epoch = 0
        dist.barrier(group=group)                      # every rank has mapped every buffer

    # ---- the exchange, in stream order ----------------------------------------------------------------
    def next_epoch(self):
        self.epoch += 1
        return self.epoch

    def qkv_ready(self, stream):
        """After the producers: publish, then wait for every rank's q/k/v slots of this exchange (one launch)."""
        _ext.sp_signal_wait(self.qkv_flag_peers, self.world, self.epoch, self.base, stream)

    def o_ready(self, stream):
        _ext.sp_signal_wait(self.o_flag_peers, self.world, self.epoch, self.base + _O_FLAG_OFF, stream)

    def attend(self, k_lens=None, mark=None):
        """q/k/v_recv (filled by the producers of this exchange) -> view of o_recv [B, s, N, 128].
        mark: optional callable(name) invoked after each phase (bench.py records CUDA events with it)."""
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self.qkv_ready(stream)
        if mark is not None:
            mark("qkv_signal_wait")
        _ext.fmha_fwd_sp(self.q_local, self.k_local, self.v_local, self.o_peers, self.world, self.rank * self.n,
                         self.N, k_lens=k_lens)
        if mark is not None:
            mark("self_attention")
        self.o_ready(stream)
        if mark is not None:
            mark("o_signal_wait")
        return self.o_local

    def close(self):
        """Unmap the peers' buffers and free this rank's (idempotent).  The caller makes sure no kernel of ANY rank
        still uses them (context teardown runs after a device synchronize + barrier when the group is still up)."""
        if self._closed:
            return
        self._closed = True
        lib = _ext.lib()
        try:
            with torch.cuda.device(self.device):
                for ptr in self._imported:
                    lib.uvb_sp_ipc_close(ptr)
                if self.base:
                    lib.uvb_sp_buffer_free(self.base)
        except Exception:       # interpreter shutdown: the driver reclaims everything anyway
            pass
        self._imported, self.base = [], 0


def context(B, s, N, device, group=None):
    """The (cached) exchange context for this shape, or None when the peer path is unavailable
    (UVB_SP_P2P=0, or CUDA IPC refused in this environment -> NCCL all-to-all is used instead)."""
    global _DISABLED
    if _DISABLED is None:
        _DISABLED = os.environ.get("UVB_SP_P2P", "1") == "0"
    if _DISABLED:
        return None
    key = (B, s, N, device.index, id(group))
    if key not in _CONTEXTS:
        try:
            # raises PeerExchangeUnavailable on EVERY rank or on none (the constructor agrees per phase)
            _CONTEXTS[key] = UlyssesP2P(B, s, N, device, group)
        except PeerExchangeUnavailable as e:   # IPC not permitted (container without shared IPC namespace, ...)
            warnings.warn(f"univid_b200: NVLink peer exchange unavailable ({e}); using NCCL all-to-all")
            _CONTEXTS[key] = None
            _DISABLED = True
    return _CONTEXTS[key]


def close_all(synchronize=True):
    """Release every exchange context of this process (buffers, IPC mappings).  Call it collectively before
    dist.destroy_process_group(); it also runs at interpreter exit."""
    ctxs = [c for c in _CONTEXTS.values() if c is not None]
    if ctxs and synchronize:
        try:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier()
        except Exception:
            pass
    for c in ctxs:
        c.close()
    _CONTEXTS.clear()


atexit.register(close_all, False)
