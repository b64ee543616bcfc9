"""Device pipeline of one LS-SPA job: reduce -> (perms -> lifts -> estimator)* -> epilogue.

This is the host-side mirror of the reference's ``ls_spa`` body
(ls_spa/ls_spa.py:157-253); all arithmetic happens in the CUDA kernels reached
through ``ops``.  ``backend`` is the only seam: the product uses ``CudaBackend``;
the CPU unit tests of the sharding / collective logic inject an oracle-backed
stand-in (tests/test_distributed_cpu.py).

Multi-GPU (one process per GPU, ``torch.distributed``):
  * rows of the reduction are sharded; the per-rank triangular factors are
    all-gathered and merged identically on every rank;
  * every super-batch of sample batches is cut into contiguous runs of batches, one
    run per rank; a rank folds its own run locally and ships ONE partial block (the
    run total); every rank folds the W totals into the (replicated) estimator state
    in rank order, computes the error estimates of its own batches from a scratch
    copy advanced by the earlier ranks' totals, and the per-batch errors are
    all-gathered, so every rank takes the same stop decision without a broadcast.
    Only a stop inside a super-batch gathers the per-batch blocks (for the replay).
"""

from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np
import torch

from . import ops
from .samplers import PermutationSource


@dataclass
class JobConfig:
    p: int
    batch_size: int
    max_samples: int | None      # None: run the source to exhaustion
    tolerance: float
    seed: int
    antithetical: bool
    estimate_errors: bool
    return_history: bool
    penultimate_check: bool = False   # the reference's extra estimate at i == max_samples - 1 (:222)


def split_batches(start: int, count: int, batch_size: int, extra_cut: int | None = None):
    """Cut samples [start, start+count) at absolute multiples of batch_size (and at
    `extra_cut`, the reference's i == max_samples-1 quirk).  Returns [(first, n), ...]."""
    out = []
    pos, end = start, start + count
    while pos < end:
        nxt = min((pos // batch_size + 1) * batch_size, end)
        if extra_cut is not None and pos < extra_cut < nxt:
            nxt = extra_cut
        out.append((pos, nxt - pos))
        pos = nxt
    return out


def contiguous_runs(nbatch: int, world: int):
    """Batches [0,nbatch) -> per-rank half-open runs of (almost) equal length."""
    per = -(-nbatch // world)
    return [(min(r * per, nbatch), min((r + 1) * per, nbatch)) for r in range(world)], per


def target_samples(p: int) -> int:
    """Samples per rank and super-batch: ~32K at p = 100 (every super-batch costs a round of collectives and one host
    synchronisation; jobs that can stop early ramp up to it, see superbatch_geometry), falling with the p^2 growth of the work per sample.
    Wide problems (the batched tile kernels, p > 152) need at least four evaluations per SM in flight."""
    t = int(32768 * (100.0 / max(p, 1)) ** 2)
    return max(592 if p > 152 else 256, min(t, 131072))


class Collective:
    """Thin wrapper over a torch.distributed process group (None = single process)."""

    def __init__(self, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.active = dist.is_available() and dist.is_initialized() and (
            group is not None or dist.get_world_size() > 1)
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """(…)-> (world, …), same shape on every rank."""
        if not self.active:
            return t.unsqueeze(0)
        flat = t.contiguous().reshape(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        self.dist.all_gather_into_tensor(out, flat, group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over the ranks (identical bits on every rank afterwards)."""
        if self.active:
            self.dist.all_reduce(t, group=self.group)
        return t

    def all_reduce_sum_int(self, v: int, device) -> int:
        if not self.active:
            return v
        t = torch.tensor([v], dtype=torch.int64, device=device)
        self.dist.all_reduce(t, group=self.group)
        return int(t.item())


# LSSPA_HOST_OVERLAP=0: no speculative problem assembly / early host preparation (A/B timing)
HOST_OVERLAP = os.environ.get("LSSPA_HOST_OVERLAP", "1") != "0"


# ---------------------------------------------------------------------------
# CUDA backend
# ---------------------------------------------------------------------------
class CudaBackend:
    name = "cuda"

    def __init__(self, device=None):
        self.device = device if device is not None else ops.require_cuda()
        self.copy_stream = None
        self.use_cholqr2 = os.environ.get("LSSPA_REDUCE", "cholqr2") != "householder"

    # -- reduction ----------------------------------------------------------
    @staticmethod
    def _host_views(X, y):
        """Host inputs as torch views.  float32 host data crosses the link as float32 (half the bytes) and is
        widened on the device -- exact, so the arithmetic is the same fp64 arithmetic on the same values; any
        other dtype is widened here."""
        Xh = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(X))
        yh = y if isinstance(y, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(y))
        if Xh.dtype not in (torch.float64, torch.float32):
            Xh = Xh.to(torch.float64)
        if yh.dtype not in (torch.float64, torch.float32):
            yh = yh.to(torch.float64)
        return Xh, yh

    @staticmethod
    def _chunk_rows(rows, p):
        return max(1, min(rows, (128 << 20) // (8 * (p + 1))))

    def start_copies(self, X, y, lo, hi, p, staging, fence):
        """Enqueue the copies of ALL chunks of the host rows [lo, hi) on the copy stream, now ->
        [(X_chunk, y_chunk, ready event, narrow X chunk or None, narrow y chunk or None)] for
        _row_chunks(prefetched=...).  The link then runs back to back with whatever the copy stream was given
        before (the other side's rows), whatever the host does in between (reading flags, drawing
        permutations).  The copy stream only ever runs DMA: float32 rows land in a float32 array of their own
        and are widened into `staging` by the consumer's stream -- a widening kernel on the copy stream would
        queue behind the persistent lift kernels that run meanwhile, and stall the link with it."""
        Xh, yh = self._host_views(X, y)
        narrow_x, narrow_y = Xh.dtype == torch.float32, yh.dtype == torch.float32
        rows = hi - lo
        chunk = self._chunk_rows(rows, p)
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
        bigX, bigy = staging
        # (allocated on the current stream like `staging`; the event orders the copies behind the allocation)
        nX = torch.empty((rows, p), dtype=torch.float32, device=self.device) if narrow_x else None
        ny = torch.empty(rows, dtype=torch.float32, device=self.device) if narrow_y else None
        if narrow_x or narrow_y:
            fence = torch.cuda.Event()
            fence.record()
        self.copy_stream.wait_event(fence)
        out = []
        with torch.cuda.stream(self.copy_stream):
            for r0 in range(lo, hi, chunk):
                r1 = min(r0 + chunk, hi)
                bx, by = bigX[r0 - lo: r1 - lo], bigy[r0 - lo: r1 - lo]
                nxc = nX[r0 - lo: r1 - lo] if narrow_x else None
                nyc = ny[r0 - lo: r1 - lo] if narrow_y else None
                (nxc if narrow_x else bx).copy_(Xh[r0:r1], non_blocking=True)
                (nyc if narrow_y else by).copy_(yh[r0:r1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
                out.append((bx, by, ready, nxc, nyc))
        return out

    def _row_chunks(self, X, y, lo, hi, p, keep=False, staging=None, fence=None, prefetched=None):
        """Yield device (X_chunk, y_chunk) covering rows [lo, hi); host inputs are streamed on a copy
        stream so that the copy of chunk i+1 overlaps the work on chunk i.  keep=False recycles two
        staging buffers (single-pass consumers); keep=True lands the chunks in one resident device
        array (the CholeskyQR2 reduction reads the rows twice).  prefetched: the copies were already
        enqueued by start_copies; only their events are awaited here."""
        if prefetched is not None:
            main = torch.cuda.current_stream()
            for bx, by, ready, nxc, nyc in prefetched:
                main.wait_event(ready)
                if nxc is not None:
                    bx.copy_(nxc)
                if nyc is not None:
                    by.copy_(nyc)
                yield bx, by, None
            return
        if isinstance(X, torch.Tensor) and X.is_cuda:
            # device-resident inputs: the kernels read raw float64 storage on THIS device, so any other
            # dtype (torch's default is float32) or device is converted here (a no-op for float64 rows)
            Xv = X[lo:hi].to(device=self.device, dtype=torch.float64)
            yv = torch.as_tensor(y)[lo:hi].to(device=self.device, dtype=torch.float64).contiguous()
            if Xv.stride(1) != 1:
                Xv = Xv.contiguous()
            yield Xv, yv, None
            return
        Xh, yh = self._host_views(X, y)
        narrow_x, narrow_y = Xh.dtype == torch.float32, yh.dtype == torch.float32
        rows = hi - lo
        chunk = self._chunk_rows(rows, p)
        if self.copy_stream is None:
            self.copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream()
        if keep:
            if staging is not None:
                bigX, bigy = staging
            else:
                bigX = torch.empty((rows, p), dtype=torch.float64, device=self.device)
                bigy = torch.empty(rows, dtype=torch.float64, device=self.device)
            bufs = None
        else:
            bufs = [(torch.empty((chunk, p), dtype=torch.float64, device=self.device),
                     torch.empty(chunk, dtype=torch.float64, device=self.device)) for _ in range(2)]
        # the buffers come from the main stream's allocator pool: whatever used that memory before
        # (e.g. the previous reduction's kernels) must finish before we copy into it.  A caller that
        # allocated the staging arrays itself passes the event recorded right after the allocation,
        # so that work it enqueued on the main stream since then does not hold the copies back.
        if fence is not None:
            self.copy_stream.wait_event(fence)
        else:
            self.copy_stream.wait_stream(main)
        free_ev = [None, None]
        tmpx = torch.empty((chunk, p), dtype=torch.float32, device=self.device) if narrow_x else None
        tmpy = torch.empty(chunk, dtype=torch.float32, device=self.device) if narrow_y else None
        for i, r0 in enumerate(range(lo, hi, chunk)):
            r1 = min(r0 + chunk, hi)
            if keep:
                bx, by = bigX[r0 - lo: r1 - lo], bigy[r0 - lo: r1 - lo]
            else:
                bx, by = bufs[i % 2][0][: r1 - r0], bufs[i % 2][1][: r1 - r0]
            with torch.cuda.stream(self.copy_stream):
                if not keep and free_ev[i % 2] is not None:
                    self.copy_stream.wait_event(free_ev[i % 2])
                if narrow_x:       # (the staging buffer is only touched on the copy stream, in order)
                    tmpx[: r1 - r0].copy_(Xh[r0:r1], non_blocking=True)
                    bx.copy_(tmpx[: r1 - r0])
                else:
                    bx.copy_(Xh[r0:r1], non_blocking=True)
                if narrow_y:
                    tmpy[: r1 - r0].copy_(yh[r0:r1], non_blocking=True)
                    by.copy_(tmpy[: r1 - r0])
                else:
                    by.copy_(yh[r0:r1], non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(self.copy_stream)
            main.wait_event(ready)
            done = torch.cuda.Event()
            yield bx, by, done
            done.record(main)
            free_ev[i % 2] = done

    # condition-number bound up to which the CholeskyQR2 reduction is used (it is as accurate as
    # Householder up to ~1e7); beyond it, or on a failed pivot, the Householder TSQR takes over
    CHOLQR2_MAX_COND = 1e6

    def alloc_staging(self, rows, p):
        """Resident device arrays for `rows` host rows plus the event after which they are free to be
        written by the copy stream."""
        bigX = torch.empty((rows, p), dtype=torch.float64, device=self.device)
        bigy = torch.empty(rows, dtype=torch.float64, device=self.device)
        fence = torch.cuda.Event()
        fence.record()
        return (bigX, bigy), fence

    def reduce_rows(self, X, y, lo, hi, p, divisor, staging=None, fence=None):
        """Rows [lo, hi) of [X|y]/divisor -> one triangular factor in slot layout (device), by the
        Householder TSQR (any shape, any conditioning)."""
        if hi - lo <= 0:
            return torch.zeros(ops.tsqr_slot(p), dtype=torch.float64, device=self.device)
        parts = []
        for Xc, yc, _ in self._row_chunks(X, y, lo, hi, p, staging=staging, fence=fence, keep=staging is not None):
            if Xc.shape[0] > 0:
                parts.append(ops.tsqr_rows(Xc, yc, divisor))
        stacked = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
        return ops.tsqr_merge(stacked, p)

    def _tsqr_side(self, coll, chunks, X, y, lo, hi, p, divisor, reg):
        """The Householder route of one side, all ranks: per-rank triangles, all-gather, ridge rows as one
        more triangle (reference :310), identical merge on every rank."""
        if chunks is not None:      # rows already resident on the device
            parts = [ops.tsqr_rows(Xc, yc, divisor) for Xc, yc in chunks if Xc.shape[0] > 0]
            if parts:
                f = ops.tsqr_merge(parts[0] if len(parts) == 1 else torch.cat(parts, 0), p)
            else:
                f = torch.zeros(ops.tsqr_slot(p), dtype=torch.float64, device=self.device)
        else:
            f = self.reduce_rows(X, y, lo, hi, p, divisor)
        g = coll.all_gather(f)
        if reg != 0.0:
            g = torch.cat([g, self.ridge(p, reg).unsqueeze(0)], 0)
        return self.merge_factors(g, p) if g.shape[0] > 1 else g[0]

    def reduce_start(self, coll, X, y, lo, hi, p, divisor, reg, staging=None, fence=None, is_train=False,
                     prefetched=None):
        """Launch half of one side of reduce_data: Gram pass over the rows [lo, hi) of this rank, all-reduce,
        factorisation.  Nothing is read back: the returned state goes to reduce_finish, and a caller may
        start the other side first so that one host synchronisation serves both.

        Gram route (default): every rank accumulates the Gram matrix of its rows with fp64 tensor tiles,
        the ranks ALL-REDUCE it (82 KB at p = 100, 8 MB at p = 1000), the ridge term is added to its
        diagonal and every rank factors the same matrix -- so the condition / pivot flags, and with them
        the decision to fall back, are identical everywhere."""
        small = ops.gram_supported(p)
        st = dict(coll=coll, X=X, y=y, lo=lo, hi=hi, p=p, divisor=divisor, reg=reg, is_train=is_train, small=small,
                  fac=None, flags=None)
        if not (self.use_cholqr2 and (small or ops.gram_big_supported(p))):
            return st
        count_rows = divisor is None                 # (small only, see reduce_problem)
        fac = (ops.CholQR2(p, 1.0 if count_rows else divisor, self.device) if small
               else ops.GramBig(p, divisor, self.device))
        chunks = []
        if hi - lo > 0:
            for Xc, yc, _ in self._row_chunks(X, y, lo, hi, p, keep=True, staging=staging, fence=fence,
                                              prefetched=prefetched):
                fac.add_chunk(Xc, yc)          # the Gram pass on this chunk overlaps the next copy
                chunks.append((Xc, yc))
        n_dev = None
        if small and count_rows:
            # [unscaled Gram | local row count] in one all-reduce; G / N on the device
            mine = torch.full((1,), float(hi - lo), dtype=torch.float64, device=self.device)
            red = coll.all_reduce_sum(torch.cat([fac.gram(), mine]))
            n_dev = red[-1:]
            slot, info = fac.factor(red[:-1] / n_dev, reg, want_gram=is_train)
        elif small:
            slot, info = fac.factor(coll.all_reduce_sum(fac.gram()), reg, want_gram=is_train)
        else:
            slot, info = fac.factor(coll.all_reduce_sum(fac.gram()), reg)
        q = p + 1
        # what the host needs from this side, as one small device tensor: [bad pivot, cond bound of the factor,
        # cond bound of its leading block (train side, p + 1 <= 112), sum of squares of the y column]
        lift_cond = (fac.lift_gram[q * q:q * q + 1] if (small and is_train and fac.lift_gram is not None)
                     else torch.full((1,), float("nan"), dtype=torch.float64, device=self.device))
        st.update(fac=fac, chunks=chunks, slot=slot,
                  flags=torch.cat([info, lift_cond, slot[q * q:q * q + 1]] + ([n_dev] if n_dev is not None else [])))
        if small and is_train and fac.lift_gram is not None and HOST_OVERLAP:
            # the train half of the problem in the lift kernels' layout, assembled NOW (device work only, behind
            # the factorisation) rather than after the host has read the flags: the common outcome keeps it
            R, c, _ = ops.split_factor(slot, p)
            st["spec_train"] = ops.TrainSide(R, c, gram=fac.lift_gram, cond=float("inf"))
        return st

    def reduce_finish(self, st, flags=None):
        """Decide half: -> (merged slot, TrainSide or None, sum of squares of y or None).  flags: the host copy
        of st['flags'] when the caller has already fetched it (together with the other side's).  One pass is kept
        when the factor is well conditioned; otherwise a second CholeskyQR pass (p + 1 <= 112, no ridge) or
        the Householder TSQR."""
        coll, p, reg, small, is_train, fac = st["coll"], st["p"], st["reg"], st["small"], st["is_train"], st["fac"]
        if fac is None:
            return self._tsqr_side(coll, None, st["X"], st["y"], st["lo"], st["hi"], p, st["divisor"], reg), None, None
        if flags is None:
            flags = st["flags"].cpu()
        vals = [float(v) for v in flags]
        bad, cond, lift_cond, ysq = vals[:4]
        if len(vals) > 4:                        # the global row count came with the flags (reduce_start)
            st["divisor"] = math.sqrt(vals[4])
            fac.scale = 1.0 / vals[4]
        slot, train = st["slot"], None
        if bad == 0 and small and is_train and fac.lift_gram is not None and cond <= self.SINGLE_PASS_COND:
            train = st.get("spec_train")
            if train is not None:
                train.set_cond(lift_cond)
            else:
                R, c, _ = ops.split_factor(slot, p)
                train = ops.TrainSide(R, c, gram=fac.lift_gram, cond=lift_cond)
        if bad == 0 and not small and is_train:
            # wide problems: the blocked factorisation reports no condition bound; the estimate the lift
            # route needs anyway (equilibrated train factor) decides
            R, c, _ = ops.split_factor(slot, p)
            train = ops.TrainSide(R, c)
            cond = train.cond_estimate
        elif bad == 0 and not small:
            cond = 0.0     # test side: everything downstream depends on R_te through R_te^T R_te, which Cholesky reproduces
        if bad == 0 and cond <= self.SINGLE_PASS_COND:
            return slot, train, ysq
        if small and bad == 0 and reg == 0.0 and cond <= self.CHOLQR2_MAX_COND:
            slot, info2 = fac.finish_second(coll.all_reduce_sum(fac.second_gram()))
            bad2, cond2 = (float(v) for v in info2.cpu())
            if bad2 == 0 and cond2 <= 10.0 * (p + 1):
                return slot, None, None
        return self._tsqr_side(coll, st["chunks"], st["X"], st["y"], st["lo"], st["hi"], p, st["divisor"], reg), None, None

    def reduce_side(self, coll, X, y, lo, hi, p, divisor, reg, staging=None, fence=None, is_train=False,
                    prefetched=None):
        """One side of reduce_data over all ranks, start to finish -> (merged slot, TrainSide or None, ysq or None)."""
        return self.reduce_finish(self.reduce_start(coll, X, y, lo, hi, p, divisor, reg, staging=staging, fence=fence,
                                                    is_train=is_train, prefetched=prefetched))

    # factor condition (equilibrated) up to which one Cholesky pass is kept
    SINGLE_PASS_COND = 1e3

    def merge_factors(self, factors, p):
        return ops.tsqr_merge(factors, p, group=max(int(factors.shape[0]), 2))

    def ridge(self, p, reg):
        return ops.ridge_factor(p, reg, self.device)

    def make_problem(self, train_slot, test_slot, p, train=None, ysq=None):
        R_tr, c_tr, _ = ops.split_factor(train_slot, p)
        R_te, c_te, ysq_dev = ops.split_factor(test_slot, p)
        if ysq is None:
            ysq = float(ysq_dev.item())
        return ops.ReducedProblem(R_tr, c_tr, R_te, c_te, ysq, train=train)

    # -- sample loop --------------------------------------------------------
    def lifts(self, prob, perms, antithetical):
        return ops.lifts(prob, perms, antithetical)

    def lifts_eliminate(self, prob, factors, perms, antithetical):
        return ops.lifts_eliminate(prob, factors, perms, antithetical)

    def make_estimator(self, cfg: JobConfig):
        return ops.Estimator(cfg.p, cfg.seed, cfg.estimate_errors, self.device)

    def prefix_means(self, rows, carry_sum, carry_count):
        out = torch.empty_like(rows)
        ops.prefix_means(rows, carry_sum, carry_count, out)
        return out

    def theta_r2(self, prob):
        return self.theta_r2_finish(self.theta_r2_start(prob))

    def theta_r2_start(self, prob):
        """Enqueue the epilogue kernel (it needs the reduced problem only, so a job launches it ahead of
        the sample loop and reads it with the final results)."""
        theta, r2 = ops.theta_r2(prob)
        return torch.cat([theta, r2.reshape(1)])

    def theta_r2_finish(self, packed):
        host = packed.cpu().numpy()
        return host[:-1].copy(), float(host[-1])

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.float64, device=self.device)

    def to_host_async(self, t):
        """Start the device-to-host copy of `t` now (pinned buffer, current stream); the returned
        callable waits for it and hands back the numpy array."""
        buf = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        buf.copy_(t, non_blocking=True)
        done = torch.cuda.Event()
        done.record()

        def wait():
            done.synchronize()
            return buf.numpy()
        return wait


# ---------------------------------------------------------------------------
# the job
# ---------------------------------------------------------------------------
def reduce_problem(backend, coll: Collective, X_train, X_test, y_train, y_test, reg, p,
                   n_train_global=None, row_sharded=False, prefactor=None, prepare=None):
    """Both tall-skinny reductions, row-sharded over the ranks of `coll`.

    With row_sharded=False every rank was handed the full arrays and reduces its own
    contiguous slice of rows; with row_sharded=True the inputs ARE the local shards."""
    def local_range(n):
        if row_sharded or coll.world == 1:
            return 0, n
        per = -(-n // coll.world)
        return min(coll.rank * per, n), min((coll.rank + 1) * per, n)

    n_tr_local = int(X_train.shape[0])
    fused = hasattr(backend, "reduce_start")     # the CUDA backend; the CPU stand-in of the unit tests merges triangles
    # Row-sharded inputs: the global row count N (the train rows are scaled by 1/sqrt(N), reference :309) is a
    # sum over the ranks.  On the Gram route it rides on the all-reduce of the Gram matrix and the scaling is
    # applied on the device, so a job does not begin with a collective and a host round trip of its own.
    defer_n = (n_train_global is None and row_sharded and coll.world > 1 and fused and prefactor is None
               and getattr(backend, "use_cholqr2", False) and ops.gram_supported(p))
    if n_train_global is None and not defer_n:
        n_train_global = (coll.all_reduce_sum_int(n_tr_local, backend.device)
                          if row_sharded else n_tr_local)
    lo_tr, hi_tr = local_range(n_tr_local)
    lo_te, hi_te = local_range(int(X_test.shape[0]))
    root_n = None if defer_n else math.sqrt(n_train_global)     # None: 1/sqrt(N) with N from the Gram all-reduce
    if not fused:
        def one_side(X, y, lo, hi, divisor, ridge):
            g = coll.all_gather(backend.reduce_rows(X, y, lo, hi, p, divisor))
            if ridge != 0.0:
                g = torch.cat([g, backend.ridge(p, ridge).unsqueeze(0)], 0)    # :310
            return backend.merge_factors(g, p) if g.shape[0] > 1 else g[0]
        # train side: rows scaled by 1/sqrt(N) (reference ls_spa/ls_spa.py:309,311), ridge rows sqrt(reg) I (:310);
        # test side unscaled, no ridge (:315)
        return backend.make_problem(one_side(X_train, y_train, lo_tr, hi_tr, root_n, reg),
                                    one_side(X_test, y_test, lo_te, hi_te, 1.0, 0.0), p)
    if prefactor is not None:
        # single process, host-resident test rows: the train factor is complete, so permutations
        # can already be drawn and factored while the test rows cross PCIe.  The staging arrays of
        # the test rows are allocated first: the copies then only wait for what was enqueued before
        # the factorisations, and no block freed during them can end up under a copy.
        # The copies of the test rows are enqueued on the copy stream right behind the train rows', before the
        # host waits for the train flags: the link never idles between the two sides.
        staging, fence = backend.alloc_staging(hi_te - lo_te, p)
        st_tr = backend.reduce_start(coll, X_train, y_train, lo_tr, hi_tr, p, root_n, reg, is_train=True)
        pre = None
        if hi_te - lo_te > 0 and not (isinstance(X_test, torch.Tensor) and X_test.is_cuda):
            pre = backend.start_copies(X_test, y_test, lo_te, hi_te, p, staging, fence)
        if prepare is not None:
            prepare()
        train_slot, train, _ = backend.reduce_finish(st_tr)
        train = prefactor.run(train_slot, p, train)
        test_slot, _, ysq = backend.reduce_side(coll, X_test, y_test, lo_te, hi_te, p, 1.0, 0.0, staging=staging, fence=fence,
                                                prefetched=pre)
    else:
        # both sides are enqueued before anything is read back: one host synchronisation for the whole reduction
        st_tr = backend.reduce_start(coll, X_train, y_train, lo_tr, hi_tr, p, root_n, reg, is_train=True)
        st_te = backend.reduce_start(coll, X_test, y_test, lo_te, hi_te, p, 1.0, 0.0)
        spec = None
        if st_tr["flags"] is not None and st_te["flags"] is not None:
            if st_tr.get("spec_train") is not None:
                # the whole problem in the lift kernels' layout, enqueued behind the two factorisations while
                # the host has not yet seen their flags (half a dozen small kernels that used to run one
                # host round trip apart, after the read)
                spec = backend.make_problem(st_tr["slot"], st_te["slot"], p, train=st_tr["spec_train"], ysq=float("nan"))
            if prepare is not None and HOST_OVERLAP:
                prepare()              # host work of the caller that needs no result of the reduction
                prepare = None
            both = torch.cat([st_tr["flags"], st_te["flags"]]).cpu()
            k_tr = int(st_tr["flags"].numel())
            f_tr, f_te = both[:k_tr], both[k_tr:]
        else:
            f_tr = f_te = None
        train_slot, train, _ = backend.reduce_finish(st_tr, f_tr)
        test_slot, _, ysq = backend.reduce_finish(st_te, f_te)
        if (spec is not None and ysq is not None and train is st_tr.get("spec_train")
                and train_slot is st_tr["slot"] and test_slot is st_te["slot"]):
            spec.finalize(ysq)
            return spec
    if prepare is not None and not (fused and prefactor is not None):
        prepare()
    return backend.make_problem(train_slot, test_slot, p, train=train, ysq=ysq)


class Prefactor:
    """Cholesky route of a host-resident job, first half ahead of time: as soon as the train factor
    exists, the permutations of the first super-batches are drawn and factored (lsspa_lifts_chol_factor
    needs the train side only).  The kernels run while the copy engine streams the test rows; the
    sample loop then only eliminates (lsspa_lifts_chol_eliminate).  table: first sample -> (perms,
    factors)."""

    MAX_BYTES = 12 << 30          # device memory spent on stored factors
    FACTOR_RATE = 18e6            # evaluations/s assumed for sizing the overlap window (packed factor kernel, p = 100)
    LINK_RATE = 50e9              # bytes/s assumed for the host link

    def __init__(self, backend, cfg: JobConfig, get_source, test_bytes: int):
        self.backend, self.cfg, self.get_source, self.test_bytes = backend, cfg, get_source, test_bytes
        self.table, self.train, self.source = {}, None, None

    def run(self, train_slot, p, train=None):
        from . import ops as _ops
        if train is None:
            R_tr, c_tr, _ = _ops.split_factor(train_slot, p)
            train = _ops.TrainSide(R_tr, c_tr)
        self.train = train
        if not (train.use_chol and _ops.split_route_supported(p)):
            return train
        cfg = self.cfg
        self.source = source = self.get_source()
        limit, bs_eff, sb_size = superbatch_geometry(cfg, 1, source.total)
        if limit is None:
            return train
        per_sample = 2 if cfg.antithetical else 1
        fd = int(_ops._lib().lsspa_lifts_chol_factor_doubles(p))
        max_bytes = self.MAX_BYTES
        if torch.cuda.is_available():
            # the test rows (staged before this runs) and the sample loop need room too.  Not
            # cudaMemGetInfo: it synchronises the device, and the host must keep running ahead here.
            dev = self.backend.device
            room = torch.cuda.get_device_properties(dev).total_memory - torch.cuda.memory_reserved(dev)
            max_bytes = min(max_bytes, max(room, 0) // 2)
        budget = min(max_bytes // (8 * fd), int(self.FACTOR_RATE * self.test_bytes / self.LINK_RATE))
        pos, index, evals = 0, 0, 0
        while pos < limit:
            want = min(sb_size(index), limit - pos)
            if evals + want * per_sample > budget:
                break
            if source.random_access:
                source.position = pos
            perms = source.take(want)
            if int(perms.shape[0]) != want:
                break                                   # stream ran dry: leave the rest to the loop
            self.table[pos] = (perms, _ops.lifts_factor(train, perms, cfg.antithetical))
            pos, index, evals = pos + want, index + 1, evals + want * per_sample
        return train


def batch_cap(p: int, estimate_errors: bool, free_bytes: int | None = None) -> int:
    """Most batches one super-batch may hold per rank.  Every batch costs one partial-moment block
    ((8 + p + p^2 + 1024 + 1024 p) doubles) and, with error estimates, (p + 1) x 1024 squared draws;
    without a cap that memory grows like 1 / batch_size (the reference's own tests use batch_size=2).
    Budget: 2 GB, or a quarter of the free device memory when that is less."""
    per_batch = 8 * (8 + p + p * p + 1024 + 1024 * p) + (8 * (p + 1) * 1024 if estimate_errors else 0)
    budget = 2 << 30
    if free_bytes is not None:
        budget = min(budget, free_bytes // 4)
    return max(1, int(budget // per_batch))


def superbatch_geometry(cfg: JobConfig, world: int, source_total, max_batches: int | None = None):
    """-> (limit, effective batch size, size(index)): how many samples super-batch `index` asks for
    (before clipping to the limit).  Deterministic, so that work can be prepared ahead of the loop.
    max_batches caps the batches of one super-batch per rank (default: batch_cap without a device)."""
    limit = cfg.max_samples
    if source_total is not None:
        limit = source_total if limit is None else min(limit, source_total)
    tgt = target_samples(cfg.p)
    # without error estimates batch boundaries are irrelevant: cut the super-batch into 1024-sample
    # batches so that the per-batch moment kernels run in parallel
    bs_eff = cfg.batch_size if cfg.estimate_errors else min(tgt, 1024)
    cap = max_batches if max_batches is not None else batch_cap(cfg.p, cfg.estimate_errors)
    g_local = max(1, min(-(-tgt // bs_eff), cap))
    # When the job can stop early the super-batches ramp up (2048 samples per rank, then 4x per
    # round up to the full size): a loose tolerance is then reached after little more than the work
    # it needs instead of after one full super-batch, at the price of two extra (pipelined) rounds.
    g_first = max(1, min(g_local, -(-2048 // bs_eff)))
    ramps = cfg.estimate_errors and cfg.tolerance > 0.0

    def size(index: int) -> int:
        g = min(g_local, g_first << min(2 * index, 30)) if ramps else g_local
        return g * world * bs_eff
    return limit, bs_eff, size


def run_samples(backend, coll: Collective, prob, source: PermutationSource, cfg: JobConfig, pre=None, est=None):
    """The estimator loop (reference ls_spa/ls_spa.py:196-236), one super-batch at a time.

    Returns (result dict, history or None, samples drawn).  result: count, mean, overall_error,
    attribution_errors, error_history (numpy)."""
    p, W, rank = cfg.p, coll.world, coll.rank
    limit, bs_eff, sb_size = superbatch_geometry(cfg, W, source.total)
    sb_index = {"next": 0}
    est = est if est is not None else backend.make_estimator(cfg)
    quirk = (cfg.max_samples - 1) if (cfg.penultimate_check and cfg.max_samples and cfg.estimate_errors) else None
    can_stop = cfg.estimate_errors and cfg.tolerance > 0.0

    hist_chunks = [] if cfg.return_history else None
    carry_sum = backend.zeros(p) if cfg.return_history else None
    err_parts, feat_last = [], None        # device tensors of per-batch errors (read at the end)
    err_hist = []                          # host copy, filled eagerly only when a stop is possible
    def launch_lifts(pos):
        """Stage A of a super-batch: permutations and lift rows of this rank's run of batches.
        Touches neither the estimator nor the host, so it can be issued one super-batch ahead."""
        size = sb_size(sb_index["next"])
        sb_index["next"] += 1
        want = size if limit is None else min(size, limit - pos)
        hit = pre.table.pop(pos, None) if pre is not None else None
        if hit is not None:
            # permutations of this super-batch were drawn and factored (train side) while the test
            # rows were still being copied: only the elimination against the test factor is left
            perms, factors = hit
            assert int(perms.shape[0]) == want, "prefactored super-batch does not match the plan"
            n_sb = want
            if source.random_access:
                source.position = pos + n_sb
            batches = split_batches(pos, n_sb, bs_eff, quirk)
            runs, per = contiguous_runs(len(batches), W)
            rows = backend.lifts_eliminate(prob, factors, perms, cfg.antithetical)
            return dict(pos=pos, n_sb=n_sb, dry=False, batches=batches, runs=runs, per=per, mine=batches,
                        my_count=n_sb, rows=rows)
        perms_all = None
        if not source.random_access:
            perms_all = source.take(want)        # sequential stream: every rank walks it
            n_sb = int(perms_all.shape[0])
        else:
            n_sb = want
        if n_sb == 0:
            return None
        batches = split_batches(pos, n_sb, bs_eff, quirk)
        runs, per = contiguous_runs(len(batches), W)
        b0, b1 = runs[rank]
        mine = batches[b0:b1]
        my_start = mine[0][0] if mine else pos
        my_count = sum(n for _, n in mine)
        if perms_all is not None:
            perms = perms_all[my_start - pos: my_start - pos + my_count]
        else:
            source.position = my_start
            perms = source.take(my_count)
            source.position = pos + n_sb
        rows = backend.lifts(prob, perms, cfg.antithetical) if my_count > 0 else backend.zeros(0, p)
        return dict(pos=pos, n_sb=n_sb, dry=perms_all is not None and n_sb < want, batches=batches, runs=runs,
                    per=per, mine=mine, my_count=my_count, rows=rows)

    pos, stopped = 0, False
    nxt = launch_lifts(0) if (limit is None or limit > 0) else None
    while nxt is not None:
        cur, nxt = nxt, None
        pos, n_sb, batches, runs, per = cur["pos"], cur["n_sb"], cur["batches"], cur["runs"], cur["per"]
        mine, my_count, rows = cur["mine"], cur["my_count"], cur["rows"]
        nb = len(batches)
        b0, b1 = runs[rank]
        desc, off = [], 0
        for first, n in mine:
            desc.append((off, n, first))
            off += n
        part = est.partials(rows, desc)
        counts = [n for _, n in batches]
        if W > 1:
            # Hierarchical merge: every rank folds its own run of batches into a scratch estimator and
            # ships the run total as ONE partial block (W small blocks on the wire instead of every
            # batch's block); the error estimates of the own batches come from a scratch copy of the
            # global state advanced by the totals of the earlier ranks; the global state absorbs the
            # W totals in rank order on every rank (replicated, deterministic).
            own_counts, nb_r = counts[b0:b1], b1 - b0
            run_counts = [sum(counts[a:b]) for a, b in runs]
            loc = est.scratch()
            if hasattr(est, "block_total"):
                own_total = est.block_total(part, nb_r)      # parallel sums: the merge is associative
            else:                                            # (CPU stand-in of the unit tests: sequential fold)
                loc.reset()
                loc.absorb(part, list(range(nb_r)), own_counts, emit=False)
                own_total = loc.export_block()
            totals = coll.all_gather(own_total.view(1, -1)).reshape(W, -1)
            loc.copy_from(est)
            if rank > 0:
                loc.absorb(totals, list(range(rank)), run_counts[:rank], emit=False)
            overall, feat = loc.absorb(part, list(range(nb_r)), own_counts, own=(0, nb_r),
                                       emit=cfg.estimate_errors)
            snap = est.snapshot() if can_stop else None
            est.absorb(totals, list(range(W)), run_counts, emit=False)
        else:
            snap = est.snapshot() if can_stop else None
            overall, feat = est.absorb(part, list(range(nb)), counts, own=(b0, b1), emit=cfg.estimate_errors)
        rows_in_order = None
        if hist_chunks is not None:
            if W > 1:
                per_rows = max(sum(n for _, n in batches[a:b]) for a, b in runs)
                padded = backend.zeros(per_rows, p)
                padded[:my_count] = rows
                allrows = coll.all_gather(padded)
                rows_in_order = torch.cat(
                    [allrows[r, :sum(n for _, n in batches[a:b])] for r, (a, b) in enumerate(runs)], 0)
            else:
                rows_in_order = rows
        keep = nb
        more = (not cur["dry"]) and (limit is None or pos + n_sb < limit)
        if cfg.estimate_errors:
            if W > 1:
                both = backend.zeros(per, p + 1)          # one exchange: [overall | per-feature] per batch
                if overall is not None:
                    both[: b1 - b0, 0] = overall
                    both[: b1 - b0, 1:] = feat
                both_all = coll.all_gather(both)
                merged = torch.cat([both_all[r, : b - a] for r, (a, b) in enumerate(runs)], 0)
                overall, feat = merged[:, 0].contiguous(), merged[:, 1:]
            if can_stop:
                # The per-batch errors start their way to the host, the lifts of the NEXT super-batch
                # are issued behind them, and only then does the host wait: the stop test (the only
                # host sync of a super-batch) is hidden behind device work.  If the test stops the
                # job, the speculative rows are dropped; they never touched the estimator.
                errs_t = (backend.to_host_async(overall) if hasattr(backend, "to_host_async")
                          else (lambda t=overall: t.cpu().numpy()))
                if more:
                    nxt = launch_lifts(pos + n_sb)
                errs = errs_t()
                hit = np.nonzero(errs < cfg.tolerance)[0]         # strict <, reference :229
                if hit.size:
                    keep = int(hit[0]) + 1
                    stopped = True
                    nxt = None
                    if keep < nb:                                 # stop inside the super-batch: replay
                        est.restore(snap)
                        if W > 1:                                 # now every batch's block is needed, in order
                            padded = part
                            if padded.shape[0] < per:
                                pad = backend.zeros(per - padded.shape[0], padded.shape[1])
                                padded = torch.cat([padded, pad], 0)
                            gathered = coll.all_gather(padded)                     # (W, per, PD)
                            slots = [r * per + i for r, (a, b) in enumerate(runs) for i in range(b - a)]
                        else:
                            gathered, slots = part, list(range(nb))
                        # (the per-feature errors are only computed for the last batch of a fold: the
                        # replay, which ends on the stop batch, supplies them)
                        _, feat_r = est.absorb(gathered, slots[:keep], counts[:keep], own=(keep - 1, keep), emit=True)
                        feat_last = feat_r[-1]
                    else:
                        feat_last = feat[keep - 1]
                else:
                    feat_last = feat[nb - 1]
                err_hist.extend(errs[:keep].tolist())
            else:
                err_parts.append(overall)
                feat_last = feat[nb - 1]
        if hist_chunks is not None:
            kept_rows = sum(counts[:keep])
            hist_chunks.append(backend.prefix_means(rows_in_order[:kept_rows].contiguous(), carry_sum, pos))
        pos += n_sb
        if more and not stopped and nxt is None:
            nxt = launch_lifts(pos)
    if hasattr(source, "check"):
        source.check()
    if hasattr(prob, "check"):
        prob.check()
    if isinstance(feat_last, torch.Tensor) and feat_last.is_cuda and hasattr(est, "state"):
        # mean, per-feature errors and (when nothing could stop the job) the error history in ONE transfer
        tail = [est.state[est.cur * p:(est.cur + 1) * p], feat_last.reshape(-1)] + (err_parts if err_parts else [])
        host = torch.cat(tail).cpu().numpy()
        res = dict(count=est.count, mean=host[:p].copy())
        feat_host = host[p:2 * p].copy()
        if err_parts:
            err_hist = host[2 * p:].tolist()
    else:
        res = est.read()
        if err_parts:
            err_hist = torch.cat(err_parts).cpu().numpy().tolist()
        feat_host = (feat_last.cpu().numpy().copy() if feat_last is not None else np.zeros(p))
    res["error_history"] = np.asarray(err_hist, dtype=np.float64)
    res["n_history"] = len(err_hist)
    res["overall_error"] = float(err_hist[-1]) if err_hist else 0.0
    res["attribution_errors"] = feat_host
    history = None
    if hist_chunks is not None:
        history = (torch.cat(hist_chunks, 0)[: res["count"]].cpu().numpy() if hist_chunks else np.zeros((0, p)))
    return res, history, pos
