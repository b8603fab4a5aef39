"""Multi-GPU sharding of the candidate set: one process per GPU, no data-path collective.

Candidates are independent, so every rank evaluates a contiguous slab of the flattened candidate index (for a
grid: a block of source-depth rows; with a tilt axis: a block of (tilt, depth) rows) with replicated mode tables.
The only exchange is the final reduction the reference does with ``np.unravel_index(np.argmax(surface))``
(ref/projects/swellex96_localization/data/collate.py:50): an all-gather of one ``(value, flat index)`` pair per
rank -- 16 bytes -- followed by a local reduce with np.argmax's first-occurrence tie rule.  Optionally the slabs
themselves are all-gathered when every rank wants the full surface.

Works over any ``torch.distributed`` backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).  On one box the
16-byte records travel without a collective library at all: :class:`PeerExchange` maps every rank's record buffer into
its peers (CUDA IPC); the record is stored into all peers over NVLink -- four 8-byte words that each carry the step's
epoch, so no fence or flag is needed -- from the last CTA of the surface kernel itself (``bogp_bartlett_grid_argext_peer``)
or by a single-block kernel (``bogp_peer_exchange``, csrc/kernels_peer.cu), a few microseconds instead of the ~25 us of an
NCCL all-gather.  ``set_pipelined`` / ``collect`` defer the wait by one step.  :func:`sharded_candidates` and
:func:`sharded_tilt_volume` are the same pattern for per-candidate environments (configs[4]) and the tilt volume
(configs[3]).  Measured: 94.5 % weak-scaling efficiency on 8 GPUs (profiles/scaling_r2.md).
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist


def slab(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced ``[lo, hi)`` row range of ``rank`` (first ``n_rows % world`` ranks get one extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def combine_extrema(vals: torch.Tensor, idxs: torch.Tensor, maximize: bool = True) -> Tuple[float, int]:
    """Reduce per-rank ``(value, global flat index)`` pairs; NaN / empty shards (index < 0) never win, ties go to
    the lowest index (np.argmax: first occurrence)."""
    best_v, best_i = float("nan"), -1
    for v, i in zip(vals.tolist(), idxs.tolist()):
        if i < 0 or v != v:
            continue
        better = best_i < 0 or (v > best_v if maximize else v < best_v) or (v == best_v and i < best_i)
        if better:
            best_v, best_i = v, int(i)
    return best_v, best_i


def allgather_extremum(local_val: float, local_idx: int, *, maximize: bool = True, device=None,
                       group: Optional[dist.ProcessGroup] = None) -> Tuple[float, int]:
    """All-gather the per-rank extremum pair and reduce locally (every rank returns the same answer)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(local_val), int(local_idx)
    world = dist.get_world_size(group)
    dev = torch.device("cpu") if device is None else torch.device(device)
    v = torch.tensor([local_val], dtype=torch.float64, device=dev)
    i = torch.tensor([local_idx], dtype=torch.int64, device=dev)
    vs = torch.empty(world, dtype=torch.float64, device=dev)
    is_ = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(vs, v, group=group) if dev.type == "cuda" else dist.all_gather(list(vs.split(1)), v, group=group)
    dist.all_gather_into_tensor(is_, i, group=group) if dev.type == "cuda" else dist.all_gather(list(is_.split(1)), i, group=group)
    return combine_extrema(vs.cpu(), is_.cpu(), maximize)


def _extremum_cuda(local: torch.Tensor, index_offset: int, maximize: bool, world: int,
                   group: Optional[dist.ProcessGroup]) -> Tuple[float, int]:
    """Device tensors: the local extremum never visits the host on its own -- (global index, value bits) is packed into one
    int64[2] record on the device, ONE all-gather carries the records and ONE host read returns them (the generic form
    costs two collectives and four host reads per step).  Used for world > 1 only: on one GPU the two host reads of the
    generic form are cheaper than the handful of small device operations that build the record (measured on the 0.05 ms
    environment kernel: 0.13 against 0.23 ms per step at N = 1, 0.27 against 0.25 ms at N = 2)."""
    dev = local.device
    if local.numel():
        flat = local.reshape(-1)
        clean = torch.where(torch.isnan(flat), torch.full_like(flat, float("-inf") if maximize else float("inf")), flat)
        li = torch.argmax(clean) if maximize else torch.argmin(clean)
        lv = flat[li].to(torch.float64)
        gi = torch.where(lv != lv, torch.full_like(li, -1), li + index_offset)
        rec = torch.stack([gi.to(torch.int64), lv.reshape(1).view(torch.int64)[0]])
    else:
        rec = torch.tensor([-1, 0], dtype=torch.int64, device=dev)
    if world > 1:
        every = torch.empty(2 * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, rec, group=group)
    else:
        every = rec
    host = every.cpu().reshape(-1, 2)
    return combine_extrema(host[:, 1].contiguous().view(torch.float64), host[:, 0], maximize)


def sharded_surface(evaluate_rows: Callable[[int, int], torch.Tensor], n_rows: int, n_cols: int, *,
                    maximize: bool = True, gather_surface: bool = False,
                    group: Optional[dist.ProcessGroup] = None):
    """Evaluate rows ``[lo, hi)`` of an ``[n_rows, n_cols]`` surface on this rank and reduce the extremum globally.

    ``evaluate_rows(lo, hi) -> tensor[hi-lo, n_cols]`` is the local hot path (the CUDA grid kernel on the GPU
    box).  Returns ``(local_slab, (value, (row, col)), full_surface_or_None)``.
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = slab(n_rows, rank, world)
    local = evaluate_rows(lo, hi)
    if local.is_cuda and not gather_surface and world > 1:
        val, idx = _extremum_cuda(local, lo * n_cols, maximize, world, group)
        pos = tuple(int(x) for x in np.unravel_index(idx, (n_rows, n_cols))) if idx >= 0 else (-1, -1)
        return local, (val, pos), None
    if local.numel():
        flat = local.reshape(-1)
        clean = torch.where(torch.isnan(flat), torch.full_like(flat, float("-inf") if maximize else float("inf")), flat)
        li = int(torch.argmax(clean) if maximize else torch.argmin(clean))
        lv = float(flat[li])
        gi = lo * n_cols + li
        if lv != lv:
            gi = -1
    else:
        lv, gi = float("nan"), -1
    dev = local.device if local.is_cuda else None
    val, idx = allgather_extremum(lv, gi, maximize=maximize, device=dev, group=group)
    full = None
    if gather_surface:
        if world == 1:
            full = local
        else:
            # slabs may differ by one row: pad to the largest, gather, trim
            rows_max = slab(n_rows, 0, world)[1]
            pad = torch.zeros(rows_max, n_cols, dtype=local.dtype, device=local.device)
            pad[: hi - lo] = local
            parts = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(parts, pad, group=group)
            full = torch.cat([p[: slab(n_rows, r, world)[1] - slab(n_rows, r, world)[0]] for r, p in enumerate(parts)])
    pos = tuple(int(x) for x in np.unravel_index(idx, (n_rows, n_cols))) if idx >= 0 else (-1, -1)
    return local, (val, pos), full


def sharded_candidates(evaluate: Callable[[int, int], torch.Tensor], n: int, *, maximize: bool = True, gather: bool = False,
                       group: Optional[dist.ProcessGroup] = None):
    """Scattered candidates and per-candidate environments (BASELINE configs[4]: 10^4 environments, each with its own mode
    set -- the process pool of ref/projects/swellex96_inv/data/bo/obj.py:28-30): rank ``r`` evaluates the candidates
    ``[lo, hi)`` of its slab with its own ``EnvBatch`` / ``bartlett_points`` call (``evaluate(lo, hi) -> values[hi - lo]``),
    nothing but the 16-byte extremum record is exchanged.  Returns ``(local_values, (value, global index), all_values |
    None)``; ``maximize=False`` for the ``cbf_ml`` objective the inversion minimises."""
    local, (val, pos), full = sharded_surface(lambda lo, hi: evaluate(lo, hi).reshape(-1, 1), n, 1, maximize=maximize,
                                              gather_surface=gather, group=group)
    return local.reshape(-1), (val, pos[0]), None if full is None else full.reshape(-1)


def sharded_tilt_volume(evaluate_tilts: Callable[[int, int], torch.Tensor], n_tilt: int, n_z: int, n_r: int, *,
                        maximize: bool = True, group: Optional[dist.ProcessGroup] = None):
    """The (tilt, depth, range) volume of BASELINE configs[3] split over the tilt axis: rank ``r`` evaluates the tilt planes
    ``[lo, hi)`` (``evaluate_tilts(lo, hi) -> [hi - lo, n_z, n_r]``, one ``bartlett_grid(r, z, tilt[lo:hi])`` call -- a unit
    of the tilt engine is a chunk of tilts of one depth, so whole planes keep its A-operand reuse).  Returns
    ``(local_planes, (value, (tilt, depth, range) index))``."""
    local, (val, pos), _ = sharded_surface(lambda lo, hi: evaluate_tilts(lo, hi).reshape(hi - lo, n_z * n_r), n_tilt, n_z * n_r,
                                           maximize=maximize, group=group)
    t, rest = pos
    zr = tuple(int(x) for x in np.unravel_index(rest, (n_z, n_r))) if t >= 0 else (-1, -1)
    return local.reshape(-1, n_z, n_r), (val, (t,) + zr)


def reduce_records(recs, maximize: bool = True) -> Tuple[float, int]:
    """``(value, global flat index)`` of the extremum over the exchanged int64[world, 2] ``(index, value bits)`` records,
    with np.argmax's first-occurrence rule across ranks (ties -> lowest global index); records with index < 0 (a rank
    whose slab was all NaN) are skipped."""
    r = np.asarray(recs.cpu() if torch.is_tensor(recs) else recs, dtype=np.int64).reshape(-1, 2)
    vals = (r[:, 1] & 0xffffffff).astype(np.uint32).view(np.float32)
    best = None
    for v, i in zip(vals, r[:, 0]):
        if i < 0 or v != v:
            continue
        if best is None or (v > best[0] if maximize else v < best[0]) or (v == best[0] and i < best[1]):
            best = (float(v), int(i))
    return best if best is not None else (0.0, -1)


class PeerExchange:
    """All-gather of one ``int64[2]`` record per rank over peer memory (one box, one process per GPU).

    ``torch.distributed`` is used once, to pass the CUDA IPC handles around; ``exchange(rec, out)`` is then a single
    kernel launch on the current stream: ``out[r] = rec of rank r`` for every rank ``r``.  All ranks must call it the
    same number of times.  Raises (``BogpError`` / ``RuntimeError``) when the GPUs cannot map each other's memory;
    callers fall back to ``dist.all_gather_into_tensor``.
    """

    def __init__(self, group: Optional[dist.ProcessGroup] = None):
        import ctypes as C

        from . import _lib
        self._lib, self._C = _lib, C
        if not dist.is_initialized():
            # one process, one GPU: the exchange degenerates to a copy of the own record (same kernels, same entry points)
            self.rank, self.world = 0, 1
            self._h = C.c_void_p()
            mine = (C.c_ubyte * 64)()
            _lib.call("bogp_peer_create", C.byref(self._h), 0, 1, C.cast(mine, C.c_void_p))
            _lib.call("bogp_peer_open", self._h, C.cast(mine, C.c_void_p))
            return
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self._h = C.c_void_p()
        mine = (C.c_ubyte * 64)()
        _lib.call("bogp_peer_create", C.byref(self._h), self.rank, self.world, C.cast(mine, C.c_void_p))
        local = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
        every = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(every, local, group=group)
        blob = bytes(every.cpu().tolist())
        buf = (C.c_ubyte * len(blob)).from_buffer_copy(blob)
        failure = None
        try:
            _lib.call("bogp_peer_open", self._h, C.cast(buf, C.c_void_p))
        except Exception as exc:      # noqa: BLE001  (reported below, on every rank, after the collective)
            failure = exc
        # one collective that is both the agreement on success and the barrier "every rank has mapped every buffer
        # before the first record is stored"; a rank that failed must still take part in it
        ok = torch.tensor([0 if failure is not None else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.close()
            raise RuntimeError(f"peer memory mapping failed on at least one rank (rank {self.rank}: {failure})")

    def exchange(self, rec: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        C = self._C
        if not (rec.is_cuda and rec.dtype == torch.int64 and rec.numel() == 2 and rec.is_contiguous()):
            raise ValueError("rec must be a contiguous CUDA int64[2] tensor")
        if not (out.is_cuda and out.dtype == torch.int64 and out.numel() == 2 * self.world and out.is_contiguous()):
            raise ValueError("out must be a contiguous CUDA int64[world, 2] tensor")
        self._lib.call("bogp_peer_exchange", self._h, C.c_void_p(rec.data_ptr()), C.c_void_p(out.data_ptr()),
                       C.c_void_p(self._lib.current_stream_ptr()))
        return out

    def set_pipelined(self, on: bool) -> None:
        """Pipelined exchange: every following exchange (stand-alone or fused into a surface call) publishes this rank's
        record of step k without waiting for the peers and returns the records of step k - 1 (index -1 on the first
        step); :meth:`collect` returns the records of the last step.  All ranks must switch at the same step."""
        self._lib.call("bogp_peer_set_pipelined", self._h, 1 if on else 0)

    def collect(self, out: torch.Tensor) -> torch.Tensor:
        """Records of the last published step (waits for the peers that have not published it yet): one kernel."""
        C = self._C
        if not (out.is_cuda and out.dtype == torch.int64 and out.numel() == 2 * self.world and out.is_contiguous()):
            raise ValueError("out must be a contiguous CUDA int64[world, 2] tensor")
        self._lib.call("bogp_peer_collect", self._h, C.c_void_p(out.data_ptr()), C.c_void_p(self._lib.current_stream_ptr()))
        return out

    def close(self):
        if self._h is not None and self._h.value:
            self._lib.lib().bogp_peer_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
