"""1-D row-partitioned SpMM with halo exchange over NVLink.

The reference has no multi-GPU path for message passing (only single-process
nn.DataParallel, HAN/train_utils/train_eval.py:46); this is new.  One process per GPU.
Rank p owns the contiguous row block [bounds[p], bounds[p+1]) of Â, X and Y (blocks balanced
by nnz).  Its local CSR is split by column into
    A_loc : columns inside its own block (re-based)       -> needs only X_local
    A_rem : columns owned by peers, remapped into a compact, de-duplicated halo buffer
Per SpMM:   push/exchange halo rows  ||  Y = A_loc·X_local   then   Y += A_rem·halo
The halo exchange has two transports:
    "p2p"  (default on NVLink): every rank writes the rows its peers need straight into the
           peers' halo buffers with one kernel of 128-bit stores over NVLink
           (gnn_halo_push_f32: pack + transfer fused, no staging), followed by a stream-ordered
           barrier; halo buffers are cudaMalloc + CUDA IPC and ping-pong between calls.
    "nccl" all_to_all_single of packed rows (the library baseline; also the gloo/CPU path of
           the planning tests).
The summation order (local columns, then remote columns, each in CSR order) is fixed, so the
result is deterministic; it differs from the single-GPU order only by fp32 re-association.

The PLAN (`build_halo_plan`) is device-agnostic host logic on plain tensors so that it is
testable with gloo on CPU; only `PartitionedSpmm` touches the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib


def balanced_bounds(rowptr_or_deg_prefix: torch.Tensor, world: int) -> List[int]:
    """Row-block boundaries with (nearly) equal nnz per block.  Input: the global int64 rowptr
    (exclusive prefix sum of degrees, length N+1)."""
    rp = rowptr_or_deg_prefix
    n = rp.numel() - 1
    total = int(rp[-1].item())
    targets = torch.tensor([total * p // world for p in range(1, world)], dtype=rp.dtype, device=rp.device)
    cuts = torch.searchsorted(rp, targets).clamp_(0, n).tolist() if world > 1 else []
    bounds = [0] + [int(c) for c in cuts] + [n]
    for i in range(1, len(bounds)):
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


@dataclass
class HaloPlan:
    rank: int
    world: int
    bounds: List[int]
    # local / remote split of this rank's rows (col ids re-based / remapped to halo slots)
    rowptr_loc: torch.Tensor
    col_loc: torch.Tensor
    val_loc: Optional[torch.Tensor]
    rowptr_rem: torch.Tensor
    col_rem: torch.Tensor
    val_rem: Optional[torch.Tensor]
    halo_ids: torch.Tensor        # sorted unique global column ids this rank needs from peers
    recv_counts: List[int]        # halo rows per owner q (halo is ordered by owner, then id)
    send_rows: torch.Tensor       # local row indices this rank must send, grouped by requester
    send_counts: List[int]        # rows per requester p
    dst_off: List[int]            # where, in requester p's halo, this rank's segment starts

    @property
    def n_local(self) -> int:
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    @property
    def n_halo(self) -> int:
        return int(self.halo_ids.numel())


def split_columns(rowptr, col, val, lo: int, hi: int):
    """Split one row block's CSR by column ownership.  Order within a row is preserved."""
    is_loc = (col >= lo) & (col < hi)
    csum = torch.zeros(col.numel() + 1, dtype=torch.int64, device=col.device)
    torch.cumsum(is_loc.to(torch.int64), 0, out=csum[1:])
    rowptr_loc = csum[rowptr]
    rowptr_rem = rowptr - rowptr_loc
    col_loc = (col[is_loc] - lo).to(torch.int32)
    rem = ~is_loc
    col_rem_global = col[rem]
    val_loc = None if val is None else val[is_loc]
    val_rem = None if val is None else val[rem]
    return rowptr_loc, col_loc, val_loc, rowptr_rem, col_rem_global, val_rem


def build_halo_plan(rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], bounds: List[int],
                    rank: int, world: int, group=None) -> HaloPlan:
    """Plan for this rank's row block (rowptr over its own rows, GLOBAL column ids).
    Collective: every rank of `group` must call it (exchanges the request lists)."""
    dev = col.device
    lo, hi = bounds[rank], bounds[rank + 1]
    rowptr_loc, col_loc, val_loc, rowptr_rem, col_rem_g, val_rem = split_columns(rowptr, col.to(torch.int64), val, lo, hi)
    halo_ids, inverse = torch.unique(col_rem_g, sorted=True, return_inverse=True)  # de-duplicated halo
    col_rem = inverse.to(torch.int32)
    edges = torch.tensor(bounds[1:-1], dtype=torch.int64, device=dev)
    owner = torch.bucketize(halo_ids, edges, right=True) if world > 1 else torch.zeros_like(halo_ids)
    recv_counts = torch.bincount(owner, minlength=world).tolist() if halo_ids.numel() else [0] * world
    recv_off = [0]
    for c in recv_counts:
        recv_off.append(recv_off[-1] + c)
    if world == 1:
        return HaloPlan(rank, world, bounds, rowptr_loc, col_loc, val_loc, rowptr_rem, col_rem, val_rem, halo_ids,
                        recv_counts, torch.zeros(0, dtype=torch.int32, device=dev), [0], [0])
    # tell every owner how many rows we want and where its segment starts in our halo
    meta_out = torch.tensor([[recv_counts[q], recv_off[q]] for q in range(world)], dtype=torch.int64, device=dev)
    meta_in = torch.empty_like(meta_out)
    dist.all_to_all_single(meta_in, meta_out, group=group)
    send_counts = meta_in[:, 0].tolist()
    dst_off = meta_in[:, 1].tolist()
    # the requested ids, as row indices local to their owner
    bounds_t = torch.tensor(bounds, dtype=torch.int64, device=dev)
    req_out = (halo_ids - bounds_t[owner]).contiguous()
    req_in = torch.empty(int(sum(send_counts)), dtype=torch.int64, device=dev)
    dist.all_to_all_single(req_in, req_out, output_split_sizes=[int(c) for c in send_counts],
                           input_split_sizes=[int(c) for c in recv_counts], group=group)
    return HaloPlan(rank, world, bounds, rowptr_loc, col_loc, val_loc, rowptr_rem, col_rem, val_rem, halo_ids,
                    [int(c) for c in recv_counts], req_in.to(torch.int32), [int(c) for c in send_counts],
                    [int(o) for o in dst_off])


class _RawCudaBuffer:
    """Expose a raw device pointer to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class PartitionedSpmm:
    """Executes Y_local = (Â·X)[own rows] for one rank.  fp32."""

    def __init__(self, plan: HaloPlan, F: int, device, group=None, transport: str = "p2p",
                 dedicated_sms: Optional[int] = None):
        """dedicated_sms: SMs the fused NVLink push gets to itself while the local-column pass runs on
        the rest (peer.cu).  None = measured default: 32 at 8 GPUs, where the exchange is the critical
        path (r01, papers100M-shaped: 36.4 ms on shared SMs -> 35.7 / 30.5 / 31.7 ms with 24 / 32 / 48
        dedicated; 16 SMs cannot feed NVLink: 47 ms); 0 (one small push CTA on every SM) below that."""
        from .graph import CSRGraph
        self.plan, self.F, self.dev, self.group = plan, int(F), torch.device(device), group
        self.transport = transport if plan.world > 1 else "none"
        self.dedicated = (32 if plan.world >= 8 else 0) if dedicated_sms is None else int(dedicated_sms)
        self.tma = False  # experimental TMA mover of peer.cu ("halo.tma"); not validated on hardware yet
        n_loc, n_halo = plan.n_local, max(plan.n_halo, 1)
        self.A_loc = CSRGraph(plan.rowptr_loc, plan.col_loc, plan.val_loc, n_loc, n_loc)
        self.A_rem = CSRGraph(plan.rowptr_rem, plan.col_rem, plan.val_rem, n_loc, n_halo)
        self.ld = (self.F + 3) // 4 * 4
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1)  # halo traffic is scheduled first
        self._ev_x = torch.cuda.Event()
        self._ev_halo = torch.cuda.Event()
        self._flag = torch.zeros(1, device=self.dev)
        self._step = 0
        self._peer_ptrs = None
        lib = _lib.load()
        if self.transport == "ce":
            # experimental (not validated on hardware yet): same IPC halo buffers as "p2p", but the rows are
            # first packed into a local send buffer and then moved by the copy engines, one copy per peer
            self._ce = True
            self.transport = "p2p"
        else:
            self._ce = False
        if self.transport == "p2p":
            # two halo buffers (ping-pong), exported to every peer through CUDA IPC
            self._own, handles = [], []
            for _ in range(2):
                p = C.c_void_p()
                h = (C.c_ubyte * 64)()
                _lib.check(lib.gnn_peer_alloc(n_halo * self.ld * 4, C.byref(p), h), "gnn_peer_alloc")
                self._own.append(p.value)
                handles.append(bytes(h))
            gathered = [None] * plan.world
            dist.all_gather_object(gathered, handles, group=group)
            self._peer_ptrs = []
            for b in range(2):
                row = []
                for q in range(plan.world):
                    if q == plan.rank:
                        row.append(self._own[b])
                    else:
                        p = C.c_void_p()
                        hb = (C.c_ubyte * 64).from_buffer_copy(gathered[q][b])
                        _lib.check(lib.gnn_peer_open(hb, C.byref(p)), "gnn_peer_open")
                        row.append(p.value)
                self._peer_ptrs.append(row)
            self.halo = [torch.as_tensor(_RawCudaBuffer(p, (n_halo, self.ld)), device=self.dev)[:, :self.F]
                         for p in self._own]
            off = [0]
            for c in plan.send_counts:
                off.append(off[-1] + c)
            self._send_off = (C.c_int64 * (plan.world + 1))(*off)
            self._dst_off = (C.c_int64 * plan.world)(*plan.dst_off)
            if self._ce:
                self._off = off
                self._sendbuf = torch.empty((max(off[-1], 1), self.ld), dtype=torch.float32, device=self.dev)
                self._copy_streams = [torch.cuda.Stream(device=self.dev, priority=-1) for _ in range(4)]
                self._ev_packed = torch.cuda.Event()
                self._ev_copied = [torch.cuda.Event() for _ in range(4)]
        elif self.transport == "nccl":
            self.halo = [torch.empty((n_halo, self.F), dtype=torch.float32, device=self.dev)]
            self._sendbuf = torch.empty((max(int(plan.send_rows.numel()), 1), self.F), dtype=torch.float32, device=self.dev)
        else:
            self.halo = [torch.zeros((n_halo, self.F), dtype=torch.float32, device=self.dev)]
        self.halo_bytes_received = plan.n_halo * self.F * 4

    # -- halo exchange ---------------------------------------------------------------------
    def _exchange(self, X: torch.Tensor) -> torch.Tensor:
        plan, lib = self.plan, _lib.load()
        if self.transport == "p2p" and self._ce:
            # pack: the push kernel with this rank's own send buffer as every "peer" (rows land in send order)
            b = self._step & 1
            cur = torch.cuda.current_stream()
            own = (C.c_void_p * plan.world)(*([self._sendbuf.data_ptr()] * plan.world))
            _lib.set_tuning("halo.dedicated_sms", 0)
            _lib.set_tuning("halo.tma", 0)
            _lib.check(lib.gnn_halo_push_f32(X.data_ptr(), X.stride(0), self.F, plan.send_rows.data_ptr(), self._send_off,
                                             own, (C.c_int64 * plan.world)(*self._off[:plan.world]), self.ld, plan.world,
                                             0, cur.cuda_stream), "gnn_halo_push_f32(pack)")
            self._ev_packed.record(cur)
            row_bytes = self.ld * 4
            for s_i in range(1, plan.world):  # rotated: rank r copies to r+1 first
                q = (plan.rank + s_i) % plan.world
                st = self._copy_streams[s_i % len(self._copy_streams)]
                st.wait_event(self._ev_packed)
                n_rows = self._off[q + 1] - self._off[q]
                _lib.check(lib.gnn_peer_copy_async(self._peer_ptrs[b][q] + plan.dst_off[q] * row_bytes,
                                                   self._sendbuf.data_ptr() + self._off[q] * row_bytes,
                                                   n_rows * row_bytes, st.cuda_stream), "gnn_peer_copy_async")
            for k, st in enumerate(self._copy_streams):
                self._ev_copied[k].record(st)
                cur.wait_event(self._ev_copied[k])
            dist.all_reduce(self._flag, group=self.group)
            return self.halo[b]
        if self.transport == "p2p":
            b = self._step & 1
            ptrs = (C.c_void_p * plan.world)(*self._peer_ptrs[b])
            _lib.set_tuning("halo.dedicated_sms", self.dedicated)
            _lib.set_tuning("halo.tma", 1 if self.tma else 0)
            _lib.check(lib.gnn_halo_push_f32(X.data_ptr(), X.stride(0), self.F, plan.send_rows.data_ptr(), self._send_off,
                                             ptrs, self._dst_off, self.ld, plan.world, (plan.rank + 1) % plan.world,
                                             torch.cuda.current_stream().cuda_stream), "gnn_halo_push_f32")
            # stream-ordered barrier: returns (on this stream) once every rank's push has completed
            dist.all_reduce(self._flag, group=self.group)
            return self.halo[b]
        if self.transport == "nccl":
            torch.index_select(X, 0, plan.send_rows.to(torch.int64), out=self._sendbuf[:plan.send_rows.numel()])
            dist.all_to_all_single(self.halo[0][:plan.n_halo], self._sendbuf[:plan.send_rows.numel()],
                                   output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts,
                                   group=self.group)
            return self.halo[0]
        return self.halo[0]

    def forward(self, X: torch.Tensor, out: Optional[torch.Tensor] = None, overlap: bool = True) -> torch.Tensor:
        from .functional import spmm_raw
        plan = self.plan
        assert X.shape == (plan.n_local, self.F) and X.dtype == torch.float32 and X.is_cuda
        main = torch.cuda.current_stream()
        if out is None:
            out = torch.empty((plan.n_local, self.F), dtype=torch.float32, device=self.dev)
        if plan.world == 1:
            return spmm_raw(self.A_loc, X, out=out)
        if overlap:
            self._ev_x.record(main)
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(self._ev_x)          # X is ready (and the previous remote pass is done)
                halo = self._exchange(X)
                self._ev_halo.record(self.comm)
            # local columns while the halo is in flight; with a dedicated push (halo.dedicated_sms) this
            # pass asks for token shared memory so the scheduler keeps it off the push's SMs
            # (the TMA mover's ring leaves 35 KB of an SM's shared memory free: ask for 40 KB then)
            excl = (40 if self.tma else 28) if (self.transport == "p2p" and self.dedicated > 0) else 0
            if excl:
                _lib.set_tuning("spmm.exclusion_smem_kb", excl)
            try:
                spmm_raw(self.A_loc, X, out=out)
            finally:
                if excl:
                    _lib.set_tuning("spmm.exclusion_smem_kb", 0)
            main.wait_event(self._ev_halo)
        else:
            halo = self._exchange(X)
            spmm_raw(self.A_loc, X, out=out)
        spmm_raw(self.A_rem, halo, out=out, accumulate=True)
        self._step += 1
        return out

    def close(self):
        lib = _lib.load()
        if self._peer_ptrs is not None:
            torch.cuda.synchronize(self.dev)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for b in range(2):
                for q, p in enumerate(self._peer_ptrs[b]):
                    if q != self.plan.rank:
                        lib.gnn_peer_close(p)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for p in self._own:
                lib.gnn_peer_free(p)
            self._peer_ptrs = None
