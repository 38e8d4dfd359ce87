"""1-D row-partitioned SpMM with a wave-pipelined halo exchange over NVLink.

The reference has no multi-GPU path for message passing (only single-process nn.DataParallel,
HAN/train_utils/train_eval.py:46); the seam this serves is `torch.spmm(adj, support)` and its autograd
(GCN/GCN.py:43) on a graph too large for one GPU.  One process per GPU.  Rank p owns the contiguous row
block [bounds[p], bounds[p+1]) of Â, X and Y (blocks balanced by nnz).

Plan (host logic on plain tensors, device-agnostic, gloo-testable — `build_halo_plan`):
  * the rank's CSR is split by column into local columns (re-based) and remote columns, the latter
    remapped into a compact, de-duplicated HALO whose slots are ordered by (owner peer, wave, id);
  * rows are cut into K nnz-balanced ROW CHUNKS; a halo row belongs to wave w if chunk w is the first
    chunk that needs it, so after waves 0..w have landed chunk w has everything it reads;
  * row classes: INTERIOR rows (no remote column) and MIXED rows.  Chunks [0, c0) are "two-pass": the
    local columns of their mixed rows are aggregated while the exchange is in flight and the remote
    columns are added after the wave lands (read-modify-write of those Y rows only); chunks [c0, K) are
    "single-pass": their mixed rows are aggregated once, from [X_local ; halo], after their wave lands —
    no read-modify-write at all.  c0 is chosen by a byte model of the overlap (or given).
Step (`PartitionedSpmm.forward`):
    comm stream:  for w in waves: push wave w to every peer (gnn_halo_push, TMA mover) ; signal flags
    main stream:  P1 = local columns of (interior rows ∪ mixed rows of two-pass chunks)
                  for w in waves: wait wave-w flags ; P2_w = the mixed rows of chunk w
There is no collective on the data path: arrival is announced per wave through monotonic flags in peer
memory (gnn_peer_signal / gnn_peer_wait), and a second flag set ("consumed") lets a single halo buffer
be reused safely.  Summation order is fixed (CSR order within each pass, local before remote), so the
result is deterministic; it differs from the single-GPU order only by fp32 re-association.

Backward (`PartitionedSpmm.backward`, dX = Âᵀ·dY for the own rows): the mirror image — the transposed
remote block turns dY into one partial row per halo slot, the (contiguous) per-owner segments travel
back over the copy engines, and the owner adds them to its local transposed product in fixed
(peer, slot) order through a pattern-only CSR over its send list: ordered summation, no atomics.

Transports: "p2p" (default: fused pack+push kernel, flags), "ce" (local pack, one copy-engine copy per
peer and wave, flags), "nccl" (all_to_all_single of packed rows — the library baseline and the path the
CPU/gloo tests replay).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from . import _lib

# Byte model of the overlap (choose_two_pass_chunks / the automatic schedule), calibrated on r02 measurements
# (papers100M-shaped, F=128 fp32): the all-to-all rate the movers sustain per direction WHILE the SpMM runs
# (2 GPUs 640 GB/s, 8 GPUs 480 GB/s; the exchange alone reaches 575-610), the gather-model rate of the SpMM alone
# (0.87-1.0 of the measured HBM peak), and the cost of one more wave (a flag wait + a short consumer launch).
NVLINK_BPS_BY_WORLD = {2: 640e9, 3: 600e9, 4: 560e9}
NVLINK_BPS_LARGE = 480e9
HBM_BPS = 5.6e12
WAVE_OVERHEAD_S = 0.3e-3
SPMM_EDGES_PER_S = 11.9e9  # F=128 rows, one B200 (1.61 G edges in 135.3 ms)
AUTO_MAX_WAVES = 8
# modelled exchange time / modelled SpMM time above which the step is exchange-bound: the mover then runs on 32 CTAs
# (8 GPUs, random graph, 31.9 GB received per rank: 59.0 ms with 32 mover CTAs, 73.0 with 64, 81.0 with 148 — more
# writers only congest the receivers' ingress; profiles/r02_dist_8gpu.jsonl)
EXCHANGE_BOUND_RATIO = 2.5


def nvlink_bps(world: int) -> float:
    return NVLINK_BPS_BY_WORLD.get(int(world), NVLINK_BPS_LARGE)


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


def select_rows(rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], rows: torch.Tensor):
    """Compact CSR over the (ascending) row subset `rows`: (rowptr', col', val')."""
    deg = rowptr[rows + 1] - rowptr[rows]
    new_rowptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=rowptr.device)
    torch.cumsum(deg, 0, out=new_rowptr[1:])
    total = int(new_rowptr[-1].item()) if rows.numel() else 0
    shift = torch.repeat_interleave(rowptr[rows] - new_rowptr[:-1], deg, output_size=total)
    idx = torch.arange(total, dtype=torch.int64, device=rowptr.device) + shift
    return new_rowptr, col[idx], None if val is None else val[idx]


@dataclass
class Consumer:
    """One SpMM launch of the step: a compact CSR over a row subset of the rank's block.
    mode "local": columns index X_local.  "remote": columns index the halo, accumulate into Y.
    "combined": columns < n_local index X_local, the others (col - n_local) the halo; overwrite."""
    mode: str
    rowptr: torch.Tensor
    col: torch.Tensor
    val: Optional[torch.Tensor]
    row_map: Optional[torch.Tensor]   # int32 compact row -> local row; None = every local row in order
    n_cols: int

    @property
    def n_rows(self) -> int:
        return int(self.rowptr.numel()) - 1

    @property
    def nnz(self) -> int:
        return int(self.col.numel())


@dataclass
class HaloPlan:
    rank: int
    world: int
    bounds: List[int]
    waves: int
    two_pass_chunks: int
    chunk_bounds: List[int]       # local row boundaries of the K row chunks
    # base local / remote split over ALL rows (the backward and the world == 1 path use it)
    rowptr_loc: torch.Tensor
    col_loc: torch.Tensor
    val_loc: Optional[torch.Tensor]
    rowptr_rem: torch.Tensor
    col_rem: torch.Tensor
    val_rem: Optional[torch.Tensor]
    halo_ids: torch.Tensor        # global column id of every halo slot, ordered by (owner, wave, id)
    recv_counts: List[int]        # halo rows per owner q
    recv_wave_counts: List[List[int]]   # [q][w]
    send_rows: torch.Tensor       # int32 local row indices this rank must send, grouped by requester, in the
    send_counts: List[int]        # requester's slot order (wave, id)
    send_wave_counts: List[List[int]]   # [p][w]
    dst_off: List[int]            # where, in requester p's halo, this rank's segment starts
    back_off: List[int]           # where, in owner q's send list, the segment requested by this rank starts
    p1: Optional[Consumer] = None
    p2: List[Optional[Consumer]] = field(default_factory=list)
    model: Dict[str, float] = field(default_factory=dict)

    @property
    def n_local(self) -> int:
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    @property
    def n_halo(self) -> int:
        return int(self.halo_ids.numel())

    def send_off(self) -> List[int]:
        off = [0]
        for c in self.send_counts:
            off.append(off[-1] + c)
        return off

    def recv_off(self) -> List[int]:
        off = [0]
        for c in self.recv_counts:
            off.append(off[-1] + c)
        return off


def choose_two_pass_chunks(K: int, F: int, elem: int, nnz_p1_base: int, rows_interior: int, m_rows, m_nnz_loc, m_nnz_rem,
                           wave_bytes, world: int = 2) -> (int, Dict[str, float]):
    """Byte model of the overlap: for every c0 in [0, K] simulate the main stream (P1, then the wave
    consumers, each gated by its wave's arrival; the SpMM runs slower while the exchange shares the HBM)
    and return the c0 with the earliest finish.  All quantities are this rank's own."""
    row_b, edge_b = F * elem, 8 + F * elem
    arrival, t = [], 0.0
    link = nvlink_bps(world)
    for w in range(K):
        t += wave_bytes[w] / link
        arrival.append(t)
    t_x = arrival[-1] if K else 0.0
    slow = max(HBM_BPS - 2 * link, 0.3 * HBM_BPS)  # exchange reads X here and lands the peers' rows here

    def run(t0, nbytes):
        if t0 >= t_x:
            return t0 + nbytes / HBM_BPS
        dt = nbytes / slow
        if t0 + dt <= t_x:
            return t0 + dt
        done = (t_x - t0) * slow
        return t_x + (nbytes - done) / HBM_BPS

    best, report = None, {}
    for c0 in range(K + 1):
        p1_rows = rows_interior + sum(m_rows[:c0])
        p1_nnz = nnz_p1_base + sum(m_nnz_loc[:c0])
        t = run(0.0, p1_nnz * edge_b + p1_rows * row_b)
        for w in range(K):
            t = max(t, arrival[w]) + WAVE_OVERHEAD_S
            if w < c0:
                t = run(t, m_nnz_rem[w] * edge_b + m_rows[w] * 2 * row_b)
            else:
                t = run(t, (m_nnz_loc[w] + m_nnz_rem[w]) * edge_b + m_rows[w] * row_b)
        report[f"c0={c0}"] = t * 1e3
        if best is None or t < best[1] - 1e-9:
            best = (c0, t)
    report["exchange_ms"] = t_x * 1e3
    nnz_all = nnz_p1_base + sum(m_nnz_loc) + sum(m_nnz_rem)
    # the whole block as one single-pass SpMM: gather-model bytes at the HBM rate, never faster than the kernel's
    # edge rate (bf16 rows halve the bytes, not the time: 137.9 ms vs 135.3 ms on one GPU)
    report["compute_ms"] = max((nnz_all * edge_b + (rows_interior + sum(m_rows)) * row_b) / HBM_BPS,
                               nnz_all / SPMM_EDGES_PER_S) * 1e3
    return best[0], report


def _auto_schedule(rowptr, col, bounds, rank, world, group, F, elem_size):
    """(K, c0) for all ranks: statistics at the finest wave count, coarsened for the smaller K (the chunk
    boundaries are nested), one all_reduce of the modelled step times."""
    dev = col.device
    lo, hi = bounds[rank], bounds[rank + 1]
    n_loc = hi - lo
    col = col.to(torch.int64)
    KM = AUTO_MAX_WAVES
    is_loc = (col >= lo) & (col < hi)
    csum = torch.zeros(col.numel() + 1, dtype=torch.int64, device=dev)
    torch.cumsum(is_loc.to(torch.int64), 0, out=csum[1:])
    loc_deg = csum[rowptr[1:]] - csum[rowptr[:-1]]
    deg = rowptr[1:] - rowptr[:-1]
    rem_deg = deg - loc_deg
    del csum
    total = int(rowptr[-1].item())
    tg = torch.tensor([total * k // KM for k in range(1, KM)], dtype=torch.int64, device=dev)
    cuts = [0] + torch.searchsorted(rowptr, tg).clamp_(0, n_loc).tolist() + [n_loc]
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    mixed = rem_deg > 0
    m_rows, m_loc, m_rem = [], [], []
    for w in range(KM):
        a, b = cuts[w], cuts[w + 1]
        mw = mixed[a:b]
        m_rows.append(int(mw.sum().item()))
        m_loc.append(int(loc_deg[a:b][mw].sum().item()))
        m_rem.append(int(rem_deg[a:b].sum().item()))
    rows_interior = n_loc - sum(m_rows)
    nnz_interior = int(loc_deg.sum().item()) - sum(m_loc)
    # halo rows per finest wave (first chunk that needs them)
    col_rem = col[~is_loc]
    del is_loc
    uniq, inverse = torch.unique(col_rem, sorted=True, return_inverse=True)
    wave_rows = [0] * KM
    if uniq.numel():
        row_of_rem = torch.repeat_interleave(torch.arange(n_loc, dtype=torch.int64, device=dev), rem_deg,
                                             output_size=int(col_rem.numel()))
        chunk_of_rem = torch.bucketize(row_of_rem, torch.tensor(cuts[1:-1], dtype=torch.int64, device=dev), right=True)
        first = torch.full((uniq.numel(),), KM, dtype=torch.int64, device=dev)
        first.scatter_reduce_(0, inverse, chunk_of_rem, reduce="amin", include_self=True)
        wave_rows = torch.bincount(first, minlength=KM).tolist()[:KM]
        del row_of_rem, chunk_of_rem, first
    del uniq, inverse, col_rem
    ks = [k for k in (1, 2, 4, 8) if k <= KM]
    times = torch.full((len(ks), KM + 1), 1e9, dtype=torch.float64, device=dev)
    for ki, K in enumerate(ks):
        gsz = KM // K
        merge = lambda v: [sum(v[w * gsz:(w + 1) * gsz]) for w in range(K)]
        _, rep_k = choose_two_pass_chunks(K, F, elem_size, nnz_interior, rows_interior, merge(m_rows), merge(m_loc),
                                          merge(m_rem), [r * F * elem_size for r in merge(wave_rows)], world)
        for c0 in range(K + 1):
            times[ki, c0] = rep_k[f"c0={c0}"]
    dist.all_reduce(times, group=group)
    best = int(torch.argmin(times).item())
    K, c0 = ks[best // (KM + 1)], best % (KM + 1)
    report = {f"K={ks[i]}": [round(float(x), 2) for x in times[i, :ks[i] + 1].tolist()] for i in range(len(ks))}
    return K, c0, report


def build_halo_plan(rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], bounds: List[int],
                    rank: int, world: int, group=None, waves: Optional[int] = 1,
                    two_pass_chunks: Optional[int] = None, F: int = 128, elem_size: int = 4,
                    consumers: bool = True) -> HaloPlan:
    """Plan for this rank's row block (rowptr over its own rows, GLOBAL column ids).
    Collective: every rank of `group` must call it with the same `waves` (exchanges the request lists).
    waves: number of row chunks / exchange waves K; None = AUTOMATIC: the byte model is evaluated for
      K in {1, 2, 4, 8} and every two-pass count on every rank, the modelled times are summed over the ranks
      and the (K, c0) with the smallest sum is used by all of them.
    two_pass_chunks: c0 in [0, waves] (None: chosen by `choose_two_pass_chunks` for feature width F)."""
    dev = col.device
    if waves is None and world > 1:
        waves, two_pass_chunks, auto_report = _auto_schedule(rowptr, col, bounds, rank, world, group, F, elem_size)
    else:
        auto_report = None
    K = max(int(waves or 1), 1) if world > 1 else 1
    lo, hi = bounds[rank], bounds[rank + 1]
    n_loc = hi - lo
    col = col.to(torch.int64)
    rowptr_loc, col_loc, val_loc, rowptr_rem, col_rem_g, val_rem = split_columns(rowptr, col, val, lo, hi)
    # K row chunks of (nearly) equal nnz
    total = int(rowptr[-1].item())
    if K > 1:
        tg = torch.tensor([total * k // K for k in range(1, K)], dtype=torch.int64, device=dev)
        cuts = torch.searchsorted(rowptr, tg).clamp_(0, n_loc).tolist()
    else:
        cuts = []
    chunk_bounds = [0] + [int(c) for c in cuts] + [n_loc]
    for i in range(1, len(chunk_bounds)):
        chunk_bounds[i] = max(chunk_bounds[i], chunk_bounds[i - 1])
    rem_deg = rowptr_rem[1:] - rowptr_rem[:-1]
    uniq, inverse = torch.unique(col_rem_g, sorted=True, return_inverse=True)  # de-duplicated halo
    edges = torch.tensor(bounds[1:-1], dtype=torch.int64, device=dev)
    owner = torch.bucketize(uniq, edges, right=True) if world > 1 else torch.zeros_like(uniq)
    if K > 1 and uniq.numel():
        row_of_rem = torch.repeat_interleave(torch.arange(n_loc, dtype=torch.int64, device=dev), rem_deg,
                                             output_size=int(col_rem_g.numel()))
        cb = torch.tensor(chunk_bounds[1:-1], dtype=torch.int64, device=dev)
        chunk_of_rem = torch.bucketize(row_of_rem, cb, right=True)
        del row_of_rem
        first_wave = torch.full((uniq.numel(),), K, dtype=torch.int64, device=dev)
        first_wave.scatter_reduce_(0, inverse, chunk_of_rem, reduce="amin", include_self=True)
        del chunk_of_rem
    else:
        first_wave = torch.zeros_like(uniq)
    key = owner * K + first_wave
    order = torch.argsort(key, stable=True)  # uniq is ascending: ties keep id order
    slot_of_uniq = torch.empty_like(order)
    slot_of_uniq[order] = torch.arange(order.numel(), dtype=torch.int64, device=dev)
    halo_ids = uniq[order]
    col_rem = slot_of_uniq[inverse].to(torch.int32)
    counts2d = (torch.bincount(key, minlength=world * K) if uniq.numel() else
                torch.zeros(world * K, dtype=torch.int64, device=dev)).view(world, K)
    recv_wave_counts = [[int(x) for x in r] for r in counts2d.tolist()]
    recv_counts = [sum(r) for r in recv_wave_counts]
    recv_off = [0]
    for c in recv_counts:
        recv_off.append(recv_off[-1] + c)
    plan_kw = dict(rank=rank, world=world, bounds=bounds, waves=K, chunk_bounds=chunk_bounds, rowptr_loc=rowptr_loc,
                   col_loc=col_loc, val_loc=val_loc, rowptr_rem=rowptr_rem, col_rem=col_rem, val_rem=val_rem,
                   halo_ids=halo_ids, recv_counts=recv_counts, recv_wave_counts=recv_wave_counts)
    if world == 1:
        return HaloPlan(two_pass_chunks=1, send_rows=torch.zeros(0, dtype=torch.int32, device=dev), send_counts=[0],
                        send_wave_counts=[[0]], dst_off=[0], back_off=[0], **plan_kw)
    # tell every owner where its segment starts in our halo and how many rows of each wave we want
    meta_out = torch.tensor([[recv_off[q]] + recv_wave_counts[q] for q in range(world)], dtype=torch.int64, device=dev)
    meta_in = torch.empty_like(meta_out)
    dist.all_to_all_single(meta_in, meta_out, group=group)
    dst_off = [int(x) for x in meta_in[:, 0].tolist()]
    send_wave_counts = [[int(x) for x in r] for r in meta_in[:, 1:].tolist()]
    send_counts = [sum(r) for r in send_wave_counts]
    # the requested ids, as row indices local to their owner, in our slot order
    bounds_t = torch.tensor(bounds, dtype=torch.int64, device=dev)
    req_out = (halo_ids - bounds_t[owner[order]]).contiguous()
    req_in = torch.empty(int(sum(send_counts)), dtype=torch.int64, device=dev)
    dist.all_to_all_single(req_in, req_out, output_split_sizes=send_counts, input_split_sizes=recv_counts, group=group)
    # backward: owner q tells requester p where p's segment starts in q's send list
    send_off = [0]
    for c in send_counts:
        send_off.append(send_off[-1] + c)
    so_out = torch.tensor(send_off[:world], dtype=torch.int64, device=dev)
    so_in = torch.empty_like(so_out)
    dist.all_to_all_single(so_in, so_out, group=group)
    plan = HaloPlan(two_pass_chunks=K, send_rows=req_in.to(torch.int32), send_counts=send_counts,
                    send_wave_counts=send_wave_counts, dst_off=dst_off, back_off=[int(x) for x in so_in.tolist()],
                    **plan_kw)
    if not consumers:
        return plan
    # ---- row classes and the consumers of the step --------------------------------------------
    mixed = rem_deg > 0
    loc_deg = rowptr_loc[1:] - rowptr_loc[:-1]
    m_rows, m_nnz_loc, m_nnz_rem = [], [], []
    for w in range(K):
        a, b = chunk_bounds[w], chunk_bounds[w + 1]
        mw = mixed[a:b]
        m_rows.append(int(mw.sum().item()))
        m_nnz_loc.append(int(loc_deg[a:b][mw].sum().item()))
        m_nnz_rem.append(int(rem_deg[a:b].sum().item()))
    rows_interior = n_loc - sum(m_rows)
    nnz_interior = int(loc_deg.sum().item()) - sum(m_nnz_loc)
    row_b = F * elem_size
    wave_bytes = [max(sum(recv_wave_counts[q][w] for q in range(world)), sum(send_wave_counts[p][w] for p in range(world)))
                  * row_b for w in range(K)]
    c0_model, report = choose_two_pass_chunks(K, F, elem_size, nnz_interior, rows_interior, m_rows, m_nnz_loc,
                                              m_nnz_rem, wave_bytes, world)
    if auto_report is not None:
        report["auto_schedule_ms_sum_over_ranks"] = auto_report
    c0 = c0_model if two_pass_chunks is None else max(0, min(int(two_pass_chunks), K))
    # exchange-bound steps (the modelled exchange outlasts the whole SpMM several times over): nothing is left to hide
    # behind the exchange, so every chunk is single-pass — the consumers park on their wave's flag and Y is written
    # once.  8 GPUs, random graph: c0 = 0 / 4 / 8 -> 60.2 / 71.4 / 65.6 ms at 32 mover CTAs (r02, r2i_rand).
    exchange_bound = world > 1 and report["exchange_ms"] >= EXCHANGE_BOUND_RATIO * max(report["compute_ms"], 1e-9)
    report["exchange_bound"] = bool(exchange_bound)
    if exchange_bound and (two_pass_chunks is None or auto_report is not None):
        c0 = 0
    plan.two_pass_chunks = c0
    plan.model = dict(report, chosen=c0, model_choice=c0_model, rows_interior=rows_interior,
                      rows_mixed=sum(m_rows), nnz_interior=nnz_interior)
    cut_row = chunk_bounds[c0]  # mixed rows at or beyond this row are single-pass
    all_rows = torch.arange(n_loc, dtype=torch.int64, device=dev)
    # P1: local columns of every row except the single-pass mixed rows
    if c0 == K:
        plan.p1 = Consumer("local", rowptr_loc, col_loc, val_loc, None, n_loc)
    else:
        keep = (~mixed) | (all_rows < cut_row)
        rows = all_rows[keep]
        rp, cc, vv = select_rows(rowptr_loc, col_loc, val_loc, rows)
        plan.p1 = Consumer("local", rp, cc, vv, rows.to(torch.int32), n_loc) if rows.numel() else None
    # combined column space of the single-pass rows: local col j -> j, remote -> n_loc + halo slot
    col_comb = None
    if c0 < K:
        is_loc = (col >= lo) & (col < hi)
        col_comb = torch.empty(col.numel(), dtype=torch.int32, device=dev)
        col_comb[is_loc] = col_loc
        col_comb[~is_loc] = col_rem + n_loc
        del is_loc
    plan.p2 = []
    for w in range(K):
        a, b = chunk_bounds[w], chunk_bounds[w + 1]
        rows = all_rows[a:b][mixed[a:b]]
        if rows.numel() == 0:
            plan.p2.append(None)
        elif w < c0:
            rp, cc, vv = select_rows(rowptr_rem, col_rem, val_rem, rows)
            plan.p2.append(Consumer("remote", rp, cc, vv, rows.to(torch.int32), max(plan.n_halo, 1)))
        else:
            rp, cc, vv = select_rows(rowptr, col_comb, val, rows)
            plan.p2.append(Consumer("combined", rp, cc, vv, rows.to(torch.int32), n_loc + max(plan.n_halo, 1)))
    return plan


def reference_step_cpu(plan: HaloPlan, X_local: torch.Tensor, halo: torch.Tensor) -> torch.Tensor:
    """The step's arithmetic with torch ops in float64 (test-side: replays P1 and the wave consumers of a
    plan exactly as `PartitionedSpmm.forward` schedules them).  `halo` = the received rows in slot order."""
    n_loc, F = plan.n_local, X_local.shape[1]
    Y = torch.zeros((n_loc, F), dtype=torch.float64)
    Xl, Hl = X_local.double(), halo.double()
    table = torch.cat([Xl, Hl], 0)

    def product(c: Consumer, src):
        rows = torch.repeat_interleave(torch.arange(c.n_rows), c.rowptr[1:] - c.rowptr[:-1])
        v = torch.ones(c.nnz, dtype=torch.float64) if c.val is None else c.val.double()
        out = torch.zeros((c.n_rows, F), dtype=torch.float64)
        out.index_add_(0, rows, src[c.col.long()] * v[:, None])
        return out

    def target(c: Consumer):
        return torch.arange(n_loc) if c.row_map is None else c.row_map.long()

    if plan.p1 is not None:
        Y[target(plan.p1)] = product(plan.p1, Xl)
    for c in plan.p2:
        if c is None:
            continue
        if c.mode == "remote":
            Y[target(c)] += product(c, Hl)
        else:
            Y[target(c)] = product(c, table)
    return Y


class _RawCudaBuffer:
    """Expose a raw device pointer to torch (zero-copy) through __cuda_array_interface__."""

    def __init__(self, ptr: int, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class _PeerBuffer:
    """One cudaMalloc allocation of this rank, exported with CUDA IPC and mapped by every peer."""

    def __init__(self, nbytes: int, rank: int, world: int, group):
        lib = _lib.load()
        self.rank, self.world, self.group = rank, world, group
        p = C.c_void_p()
        h = (C.c_ubyte * 64)()
        _lib.check(lib.gnn_peer_alloc(max(int(nbytes), 256), C.byref(p), h), "gnn_peer_alloc")
        self.own = p.value
        gathered = [None] * world
        dist.all_gather_object(gathered, bytes(h), group=group)
        self.ptrs = []
        for q in range(world):
            if q == rank:
                self.ptrs.append(self.own)
            else:
                pq = C.c_void_p()
                hb = (C.c_ubyte * 64).from_buffer_copy(gathered[q])
                _lib.check(lib.gnn_peer_open(hb, C.byref(pq)), "gnn_peer_open")
                self.ptrs.append(pq.value)

    def close(self):
        lib = _lib.load()
        if self.ptrs is None:
            return
        for q, p in enumerate(self.ptrs):
            if q != self.rank:
                lib.gnn_peer_close(p)
        if dist.is_initialized():
            dist.barrier(group=self.group)
        lib.gnn_peer_free(self.own)
        self.ptrs = None


_TORCH_TYPESTR = {torch.float32: "<f4", torch.bfloat16: "<u2"}


class PartitionedSpmm:
    """Executes Y_local = (Â·X)[own rows] (and dX_local = (Âᵀ·dY)[own rows]) for one rank.

    Every scheduling choice is a constructor argument and is passed to the C ABI per call
    (gnn_halo_opts / gnn_spmm_opts): nothing is routed through process-global knobs.
      transport      "p2p" | "ce" | "nccl"
      mover          "tma" | "vector" | "auto"  (p2p only)
      mover_ctas     CTAs of the push kernel (0 = one per SM; -1 = measured default: 4-warp movers on 32 SMs when
                     the step is single-pass / exchange-dominated, on 48 SMs otherwise — r02 sweeps at 2 and 8 GPUs:
                     more mover CTAs take SpMM occupancy without raising the all-to-all rate, fewer starve it)
      mover_warps    warps per push CTA (0 = default: TMA 1, vector 8; -1 = measured default 4)
      dedicated_sms  > 0: vector mover on that many SMs of its own (each CTA claims 200 KB of shared memory
                     and the concurrent P1 asks for 28 KB so that the block scheduler keeps them apart)
      timeout_ms     bound of every flag wait (a lost peer must not hang the GPU)
      interleave     fused-signal mover: > 0 cuts every peer segment of a wave into pieces of that many ring stages
                     and deals the pieces round-robin over the peers, so a sender feeds all its peers at once
                     (uniform ingress load on every receiver); 0 = one peer after the other, rotated by rank
      fused_signal   p2p + TMA mover: all waves of a step in ONE launch, the arrival flags raised from inside
                     the kernel by the last warp that finishes a wave (gnn_halo_push_waves); False = one
                     mover launch + one signal launch per wave
    """

    def __init__(self, plan: HaloPlan, F: int, device, group=None, transport: str = "p2p",
                 dtype: torch.dtype = torch.float32, mover: str = "auto", mover_ctas: int = -1, mover_warps: int = -1,
                 dedicated_sms: int = 0, timeout_ms: int = 20000, fused_signal: bool = True, interleave: int = 0):
        from .graph import CSRGraph
        if dtype not in _TORCH_TYPESTR:
            raise _lib.GnnError(f"PartitionedSpmm: unsupported dtype {dtype}")
        self.plan, self.F, self.dev, self.group, self.dtype = plan, int(F), torch.device(device), group, dtype
        self.elem = 4 if dtype == torch.float32 else 2
        self.transport = transport if plan.world > 1 else "none"
        if self.transport not in ("p2p", "ce", "nccl", "none"):
            raise ValueError(f"unknown transport {transport!r}")
        self.mover = {"auto": 0, "vector": 1, "tma": 2}[mover]
        if int(mover_ctas) < 0:
            model = getattr(plan, "model", None) or {}
            exchange_bound = model.get("exchange_ms", 0.0) >= EXCHANGE_BOUND_RATIO * max(model.get("compute_ms", 0.0), 1e-9)
            mover_ctas = 32 if (plan.two_pass_chunks == 0 or exchange_bound) else 48
        if int(mover_warps) < 0:
            mover_warps = 4
        self.mover_ctas, self.mover_warps, self.dedicated = int(mover_ctas), int(mover_warps), int(dedicated_sms)
        self.timeout_ms = int(timeout_ms)
        self.interleave = int(interleave)
        n_loc, n_halo = plan.n_local, max(plan.n_halo, 1)
        per16 = 16 // self.elem
        # halo rows are contiguous (ld == F) when a row is a 16-byte multiple: one bulk store per ring stage
        self.ld = self.F if self.F % per16 == 0 else (self.F + per16 - 1) // per16 * per16
        self.A_loc = CSRGraph(plan.rowptr_loc, plan.col_loc, plan.val_loc, n_loc, n_loc)
        self._cons = {}
        self._wave_table = None
        self._t_loc = self._t_rem = self._R = None
        self._fstep = self._bstep = 0
        self._bufs: List[_PeerBuffer] = []
        self._status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        if plan.world == 1:
            return
        self.comm = torch.cuda.Stream(device=self.dev, priority=-1)  # halo traffic is scheduled first
        self._ev_x = torch.cuda.Event()
        self._ev_halo = torch.cuda.Event()
        self._send_off = plan.send_off()
        self._recv_off = plan.recv_off()
        self._wave_cum_send = [[sum(r[:w]) for w in range(plan.waves + 1)] for r in plan.send_wave_counts]
        if self.transport in ("p2p", "ce"):
            self._halo_buf = _PeerBuffer(n_halo * self.ld * self.elem, plan.rank, plan.world, group)
            self._flags = _PeerBuffer(4 * 32 * 4, plan.rank, plan.world, group)  # arrive/consumed x fwd/bwd
            self._bufs += [self._halo_buf, self._flags]
            self.halo = self._view(self._halo_buf.own, n_halo, self.ld)[:, :self.F]
            self._wave_table = None
            R = _lib.load().gnn_halo_rows_per_stage(self.F * self.elem)
            if (self.transport == "p2p" and fused_signal and self.mover != 1 and self.dedicated == 0 and R > 0
                    and self.ld == self.F):
                self._wave_table = self._build_wave_table(R)
            if self.transport == "ce":
                self._sendbuf = torch.empty((max(self._send_off[-1], 1), self.ld), dtype=dtype, device=self.dev)
                self._copy_streams = [torch.cuda.Stream(device=self.dev, priority=-1) for _ in range(4)]
                self._ev_packed = torch.cuda.Event()
                self._ev_copied = [torch.cuda.Event() for _ in range(4)]
        else:
            self.halo = torch.empty((n_halo, self.F), dtype=dtype, device=self.dev)
            self._sendbuf = torch.empty((max(int(plan.send_rows.numel()), 1), self.F), dtype=dtype, device=self.dev)
        self.halo_bytes_received = plan.n_halo * self.F * self.elem

    def _build_wave_table(self, R: int):
        """Device segment table of gnn_halo_push_waves: segments sorted by wave, peers in rotated order."""
        plan, W = self.plan, self.plan.world
        row_bytes = self.ld * self.elem
        rows, chunk = [], 0
        piece = self.interleave * R  # rows per interleaved piece (0: whole segments, one peer after the other)
        for w in range(plan.waves):
            segs = []
            for s_i in range(1, W):
                q = (plan.rank + s_i) % W
                n = plan.send_wave_counts[q][w]
                if n:
                    b = self._wave_cum_send[q][w]
                    segs.append((self._send_off[q] + b, n, self._halo_buf.ptrs[q] + (plan.dst_off[q] + b) * row_bytes))
            if piece > 0:
                pieces, k = [], 0
                while any(k * piece < n for _, n, _ in segs):
                    for src, n, dst in segs:  # round-robin over the peers, rotated order kept within a round
                        if k * piece < n:
                            m = min(piece, n - k * piece)
                            pieces.append((src + k * piece, m, dst + k * piece * row_bytes))
                    k += 1
                segs = pieces
            for src, n, dst in segs:
                rows.append([src, n, dst, chunk, w])
                chunk += (n + R - 1) // R
        table = torch.tensor(rows if rows else [[0, 0, 0, 0, 0]], dtype=torch.int64, device=self.dev)
        done = torch.zeros(max(plan.waves, 1), dtype=torch.int32, device=self.dev)
        return table, len(rows), chunk, done

    # -- helpers ---------------------------------------------------------------------------
    def _view(self, ptr: int, rows: int, ld: int) -> torch.Tensor:
        t = torch.as_tensor(_RawCudaBuffer(ptr, (rows, ld), _TORCH_TYPESTR[self.dtype]), device=self.dev)
        return t.view(self.dtype) if self.dtype != torch.float32 else t

    def _flag_ptrs(self, which: int):
        """Device pointers of flag array `which` (0 arrive_fwd, 1 consumed_fwd, 2 arrive_bwd, 3 consumed_bwd)
        on every rank."""
        return (C.c_void_p * self.plan.world)(*[p + which * 128 for p in self._flags.ptrs])

    def _signal(self, which: int, value: int, stream):
        _lib.check(_lib.load().gnn_peer_signal(self._flag_ptrs(which), self.plan.world, self.plan.rank, self.plan.rank,
                                               value & 0xFFFFFFFF, stream.cuda_stream), "gnn_peer_signal")

    def _wait(self, which: int, value: int, stream):
        _lib.check(_lib.load().gnn_peer_wait(self._flags.own + which * 128, self.plan.world, self.plan.rank,
                                             value & 0xFFFFFFFF, self._status.data_ptr(), self.timeout_ms,
                                             stream.cuda_stream), "gnn_peer_wait")

    def check_status(self):
        """Raise if a flag wait of an earlier step timed out (synchronises)."""
        s = int(self._status.item())
        if s:
            raise _lib.GnnError(f"rank {self.plan.rank}: timed out waiting for the halo flag of peer {s - 1}")

    def _graph(self, c: Consumer):
        from .graph import CSRGraph
        g = self._cons.get(id(c))
        if g is None:
            g = CSRGraph(c.rowptr.to(self.dev), c.col.to(self.dev), None if c.val is None else c.val.to(self.dev),
                         c.n_rows, c.n_cols)
            g._row_map = None if c.row_map is None else c.row_map.to(self.dev).contiguous()
            self._cons[id(c)] = g
        return g

    def _push(self, X: torch.Tensor, peer_ptrs, seg_begin, seg_rows, dst_row, ld_dst, send_rows, stream,
              force_vector_shared: bool = False):
        plan, lib = self.plan, _lib.load()
        W = plan.world
        o = _lib.HaloOpts()
        o.struct_size = C.sizeof(_lib.HaloOpts)
        o.first_peer = (plan.rank + 1) % W
        if force_vector_shared:
            o.mover, o.ctas, o.warps_per_cta = 1, 0, 8
        elif self.dedicated > 0:
            o.mover, o.ctas, o.warps_per_cta, o.claim_smem_bytes = 1, self.dedicated, 32, 200 * 1024
        else:
            o.mover, o.ctas, o.warps_per_cta = self.mover, self.mover_ctas, self.mover_warps
        _lib.check(lib.gnn_halo_push(X.data_ptr(), X.stride(0), self.F, self.elem,
                                     None if send_rows is None else send_rows.data_ptr(),
                                     (C.c_int64 * W)(*seg_begin), (C.c_int64 * W)(*seg_rows),
                                     (C.c_void_p * W)(*peer_ptrs), (C.c_int64 * W)(*dst_row), ld_dst, W, C.byref(o),
                                     stream.cuda_stream), "gnn_halo_push")

    # -- forward ---------------------------------------------------------------------------
    def _exchange_wave(self, X: torch.Tensor, w: int, stream):
        """Push wave w of every peer's segment (p2p) / copy it (ce) on `stream`."""
        plan = self.plan
        W = plan.world
        seg_begin = [self._send_off[p] + self._wave_cum_send[p][w] for p in range(W)]
        seg_rows = [plan.send_wave_counts[p][w] for p in range(W)]
        dst_row = [plan.dst_off[p] + self._wave_cum_send[p][w] for p in range(W)]
        if self.transport == "p2p":
            self._push(X, self._halo_buf.ptrs, seg_begin, seg_rows, dst_row, self.ld, plan.send_rows, stream)
            return
        lib = _lib.load()
        row_bytes = self.ld * self.elem
        self._ev_packed.record(stream)
        for s_i in range(1, W):  # rotated: rank r copies to r+1 first
            q = (plan.rank + s_i) % W
            if seg_rows[q] == 0:
                continue
            st = self._copy_streams[s_i % len(self._copy_streams)]
            st.wait_event(self._ev_packed)
            _lib.check(lib.gnn_peer_copy_async(self._halo_buf.ptrs[q] + dst_row[q] * row_bytes,
                                               self._sendbuf.data_ptr() + seg_begin[q] * row_bytes,
                                               seg_rows[q] * row_bytes, st.cuda_stream), "gnn_peer_copy_async")
        for k, st in enumerate(self._copy_streams):
            self._ev_copied[k].record(st)
            stream.wait_event(self._ev_copied[k])

    def _exchange(self, X: torch.Tensor, stream):
        """Whole exchange of one step on `stream` (waves in order, flags after every wave)."""
        plan = self.plan
        K, s = plan.waves, self._fstep
        if self.transport == "nccl":
            n_send = int(plan.send_rows.numel())
            torch.index_select(X, 0, plan.send_rows.to(torch.int64), out=self._sendbuf[:n_send])
            dist.all_to_all_single(self.halo[:plan.n_halo], self._sendbuf[:n_send],
                                   output_split_sizes=plan.recv_counts, input_split_sizes=plan.send_counts,
                                   group=self.group)
            return
        if s > 0:
            self._wait(1, s, stream)  # every peer has consumed the halo of the previous step: safe to overwrite
        if self.transport == "ce":
            # pack: the push kernel with this rank's own send buffer as every "peer" (rows land in send order)
            W = plan.world
            self._push(X, [self._sendbuf.data_ptr()] * W, self._send_off[:W], plan.send_counts, self._send_off[:W],
                       self.ld, plan.send_rows, stream, force_vector_shared=True)
        if self._wave_table is not None:
            table, n_segs, n_chunks, done = self._wave_table
            o = _lib.HaloOpts()
            o.struct_size = C.sizeof(_lib.HaloOpts)
            o.ctas, o.warps_per_cta = self.mover_ctas, self.mover_warps
            _lib.check(_lib.load().gnn_halo_push_waves(
                X.data_ptr(), X.stride(0), self.F, self.elem, plan.send_rows.data_ptr(), table.data_ptr(), n_segs, K,
                n_chunks, done.data_ptr(), self._flag_ptrs(0), plan.world, plan.rank, (s * K) & 0xFFFFFFFF, C.byref(o),
                stream.cuda_stream), "gnn_halo_push_waves")
            return
        for w in range(K):
            self._exchange_wave(X, w, stream)
            self._signal(0, s * K + w + 1, stream)

    def _run(self, c: Consumer, X, halo, out, excl: int = 0):
        from .functional import spmm_ex
        g = self._graph(c)
        if c.mode == "local":
            spmm_ex(g, X, out, row_map=g._row_map, exclusion_smem=excl)
        elif c.mode == "remote":
            spmm_ex(g, halo, out, row_map=g._row_map, accumulate=True)
        else:
            spmm_ex(g, X, out, row_map=g._row_map, X2=halo, split=self.plan.n_local)

    def forward(self, X: torch.Tensor, out: Optional[torch.Tensor] = None, overlap: bool = True) -> torch.Tensor:
        from .functional import spmm_raw
        plan = self.plan
        assert X.shape == (plan.n_local, self.F) and X.dtype == self.dtype and X.is_cuda
        main = torch.cuda.current_stream()
        if out is None:
            out = torch.empty((plan.n_local, self.F), dtype=self.dtype, device=self.dev)
        if plan.world == 1:
            return spmm_raw(self.A_loc, X, out=out)
        K, s = plan.waves, self._fstep
        flags = self.transport in ("p2p", "ce")
        comm = self.comm if overlap else main
        if overlap:
            self._ev_x.record(main)
            self.comm.wait_event(self._ev_x)  # X is ready (and everything earlier on the main stream is done)
        with torch.cuda.stream(comm):
            self._exchange(X, comm)
            if not flags:
                self._ev_halo.record(comm)
        excl = 28 * 1024 if (self.transport == "p2p" and self.dedicated > 0 and overlap) else 0
        if plan.p1 is not None:
            self._run(plan.p1, X, self.halo, out, excl)
        if not flags and overlap:
            main.wait_event(self._ev_halo)
        for w in range(K):
            if flags:
                self._wait(0, s * K + w + 1, main)
            if plan.p2[w] is not None:
                self._run(plan.p2[w], X, self.halo, out)
        if flags:
            self._signal(1, s + 1, main)  # this rank is done reading its halo
        self._fstep += 1
        return out

    def exchange_only(self, X: torch.Tensor):
        """One halo exchange with no aggregation (measurement: the transport alone), flag protocol kept."""
        plan = self.plan
        if plan.world == 1:
            return
        main = torch.cuda.current_stream()
        K, s = plan.waves, self._fstep
        self._exchange(X, main)
        if self.transport in ("p2p", "ce"):
            self._wait(0, s * K + K, main)
            self._signal(1, s + 1, main)
        self._fstep += 1

    # -- backward --------------------------------------------------------------------------
    def _backward_setup(self):
        from .graph import CSRGraph, index_block_transpose
        plan = self.plan
        n_loc, n_halo = plan.n_local, max(plan.n_halo, 1)
        self._t_loc = self.A_loc.transpose()
        if plan.world == 1:
            return
        A_rem = CSRGraph(plan.rowptr_rem, plan.col_rem, plan.val_rem, n_loc, n_halo)
        self._t_rem = A_rem.transpose()          # [n_halo, n_loc]: one partial row per halo slot
        self._t_rem._t = None
        del A_rem
        n_send = int(plan.send_rows.numel())
        # R: local row r <- the positions k of the send list with send_rows[k] == r, ascending k
        rowptr_t, pos_t = index_block_transpose(plan.send_rows, n_loc)
        deg = rowptr_t[1:] - rowptr_t[:-1]
        rows = torch.nonzero(deg > 0).flatten()
        rp, cc, _ = select_rows(rowptr_t, pos_t[:max(n_send, 0)], None, rows)
        self._R = CSRGraph(rp, cc.to(torch.int32), None, int(rows.numel()), max(n_send, 1))
        self._R._row_map = rows.to(torch.int32).contiguous()
        self._partial = torch.empty((n_halo, self.ld), dtype=self.dtype, device=self.dev)[:, :self.F]
        if self.transport in ("p2p", "ce"):
            self._back_buf = _PeerBuffer(max(n_send, 1) * self.ld * self.elem, plan.rank, plan.world, self.group)
            self._bufs.append(self._back_buf)
            self.back = self._view(self._back_buf.own, max(n_send, 1), self.ld)[:, :self.F]
            self._bcopy_streams = [torch.cuda.Stream(device=self.dev, priority=-1) for _ in range(4)]
            self._ev_partial = torch.cuda.Event()
            self._ev_bcopied = [torch.cuda.Event() for _ in range(4)]
        else:
            self.back = torch.empty((max(n_send, 1), self.F), dtype=self.dtype, device=self.dev)

    def backward(self, dY: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """dX_local = (Âᵀ·dY)[own rows]: the autograd of GCN/GCN.py:43 on the partitioned graph.
        Remote partials first (they must travel), the local transposed product while they do, then the
        partials this rank's rows received are added in fixed (peer, slot) order."""
        from .functional import spmm_ex, spmm_raw
        plan, lib = self.plan, _lib.load()
        assert dY.shape == (plan.n_local, self.F) and dY.dtype == self.dtype and dY.is_cuda
        if self._t_loc is None:
            self._backward_setup()
        main = torch.cuda.current_stream()
        if out is None:
            out = torch.empty((plan.n_local, self.F), dtype=self.dtype, device=self.dev)
        if plan.world == 1:
            return spmm_raw(self._t_loc, dY, out=out)
        s, W = self._bstep, plan.world
        spmm_raw(self._t_rem, dY, out=self._partial)           # partial[h] = Σ_i Â[i, halo h] dY[i]
        if self.transport == "nccl":
            dist.all_to_all_single(self.back[:int(plan.send_rows.numel())], self._partial[:plan.n_halo].contiguous(),
                                   output_split_sizes=plan.send_counts, input_split_sizes=plan.recv_counts,
                                   group=self.group)
            spmm_raw(self._t_loc, dY, out=out)
        else:
            self._ev_partial.record(main)
            comm = self.comm
            comm.wait_event(self._ev_partial)
            with torch.cuda.stream(comm):
                if s > 0:
                    self._wait(3, s, comm)  # every owner has consumed the partials of the previous step
                row_bytes = self.ld * self.elem
                base = self._partial.data_ptr()
                for s_i in range(1, W):
                    q = (plan.rank + s_i) % W
                    if plan.recv_counts[q] == 0:
                        continue
                    st = self._bcopy_streams[s_i % len(self._bcopy_streams)]
                    st.wait_event(self._ev_partial)
                    if s > 0:
                        st.wait_stream(comm)
                    _lib.check(lib.gnn_peer_copy_async(self._back_buf.ptrs[q] + plan.back_off[q] * row_bytes,
                                                       base + self._recv_off[q] * row_bytes,
                                                       plan.recv_counts[q] * row_bytes, st.cuda_stream),
                               "gnn_peer_copy_async")
                for k, st in enumerate(self._bcopy_streams):
                    self._ev_bcopied[k].record(st)
                    comm.wait_event(self._ev_bcopied[k])
                self._signal(2, s + 1, comm)
            spmm_raw(self._t_loc, dY, out=out)                 # local transposed product, under the copies
            self._wait(2, s + 1, main)
        if self._R.n_rows > 0:
            spmm_ex(self._R, self.back, out, row_map=self._R._row_map, accumulate=True)
        if self.transport != "nccl":
            self._signal(3, s + 1, main)
        self._bstep += 1
        return out

    def close(self):
        if self._bufs:
            torch.cuda.synchronize(self.dev)
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for b in self._bufs:
                b.close()
            self._bufs = []
