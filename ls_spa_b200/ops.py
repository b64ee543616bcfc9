"""Tensor-level wrappers of the C ABI (include/lsspa.h).

PyTorch is used for device memory, streams and (in ``dist.py``) process groups
only; every computation below is one or more launches of the hand-written
sm_100a kernels in ``ls_spa_b200/csrc``.  All functions require CUDA tensors and
raise ``LsSpaCudaError`` otherwise -- there is no CPU path.
"""

from __future__ import annotations

import math
import os

import torch

from . import _cabi
from ._cabi import LsSpaCudaError, check

ERR_DRAWS = 1024

# number of kernels of libls_spa_b200.so launched through this module (bench.py reports it)
LAUNCHES = 0
# when set to a list, every lifts() call appends (start_event, end_event, permutations)
LIFT_TRACE = None
# |R_tr|_F |R_tr^-1|_F above which the per-permutation core keeps to Householder reflections:
# the Cholesky route loses eps * cond^2, 1e3 keeps that at ~1e-10 of the lift scale.
CHOL_COND_LIMIT = 1e3
# route taken by the most recent lifts() call: 'cholesky' or 'householder' (bench.py reports it)
LIFT_ROUTE = None


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def _lib():
    return _cabi.load()


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise LsSpaCudaError("ls_spa_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t: torch.Tensor | None) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _dev_f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise LsSpaCudaError(f"{name} must be a CUDA tensor")
    if t.dtype != torch.float64:
        t = t.to(torch.float64)
    return t


# ---------------------------------------------------------------------------
# reduction
# ---------------------------------------------------------------------------
def tsqr_slot(p: int) -> int:
    return int(_lib().lsspa_tsqr_slot_doubles(p))


def tsqr_rows(X: torch.Tensor, y: torch.Tensor, divisor: float = 1.0) -> torch.Tensor:
    """Per-CTA triangular factors of [X | y] / divisor -> (nparts, slot) float64."""
    X = _dev_f64(X, "X")
    y = _dev_f64(y, "y").contiguous()
    if X.dim() != 2 or y.dim() != 1 or X.shape[0] != y.shape[0]:
        raise LsSpaCudaError("tsqr_rows: X must be (n, p) and y (n,)")
    if X.stride(1) != 1:
        X = X.contiguous()
    n, p = X.shape
    lib = _lib()
    nparts = lib.lsspa_tsqr_num_parts(p, n)
    parts = torch.empty((nparts, tsqr_slot(p)), dtype=torch.float64, device=X.device)
    check(lib.lsspa_tsqr_rows(X.data_ptr(), X.stride(0), y.data_ptr(), n, p, float(divisor),
                              parts.data_ptr(), nparts, _stream()), "lsspa_tsqr_rows")
    _count(1)
    return parts


def tsqr_merge(parts: torch.Tensor, p: int, group: int = 8) -> torch.Tensor:
    """Tree-merge stacked triangular factors (count, slot) down to one (slot,) factor."""
    lib = _lib()
    parts = parts.contiguous()
    while True:
        count = parts.shape[0]
        nout = (count + group - 1) // group
        out = torch.empty((nout, parts.shape[1]), dtype=torch.float64, device=parts.device)
        check(lib.lsspa_tsqr_merge(parts.data_ptr(), count, group, p, out.data_ptr(), _stream()),
              "lsspa_tsqr_merge")
        _count(1)
        parts = out
        if nout == 1:
            return parts[0]


def gram_supported(p: int) -> bool:
    return bool(_lib().lsspa_gram_supported(p))


class CholQR2:
    """CholeskyQR(2) of [X | y] / divisor over device row chunks (csrc/gram.cu), p + 1 <= 112.

    add_chunk() launches the Gram pass on a chunk as soon as it is resident (so it overlaps the copy
    of the next one); gram() sums the partial Gram matrices of this process (a multi-GPU job
    all-reduces that tensor: 82 KB at p = 100); factor(G, reg) adds the ridge term reg I (reference
    ls_spa/ls_spa.py:310) to the diagonal, factors, and returns (slot, info): slot has the layout of
    tsqr_merge's result, info (device) = [bad pivot, cond bound].  second_gram() / finish_second()
    are the second pass of CholeskyQR2 for factors that are not well conditioned."""

    # |R1|_F |R1^-1|_F of [X | y] up to which one Cholesky pass is kept as the factor
    SINGLE_PASS_COND = 1e3

    def __init__(self, p: int, divisor: float, device=None):
        self.p, self.scale = p, 1.0 / (float(divisor) ** 2)
        self.chunks, self.parts1 = [], []
        self.n2 = int(_lib().lsspa_gram_slot_doubles(p))
        self.device = device

    def _rows(self, Xc, yc, rinv):
        lib = _lib()
        n = Xc.shape[0]
        nparts = lib.lsspa_gram_num_parts(self.p, n, 1 if rinv is not None else 0)
        buf = torch.empty((nparts, self.n2), dtype=torch.float64, device=Xc.device)
        check(lib.lsspa_gram_rows(Xc.data_ptr(), Xc.stride(0), yc.data_ptr(), n, self.p, _ptr(rinv),
                                  buf.data_ptr(), nparts, _stream()), "lsspa_gram_rows")
        _count(1)
        return buf

    def _finish(self, parts, scale):
        G = torch.empty(self.n2, dtype=torch.float64, device=parts.device)
        check(_lib().lsspa_gram_finish(parts.data_ptr(), parts.shape[0], self.p, scale, G.data_ptr(),
                                       _stream()), "lsspa_gram_finish")
        _count(1)
        return G

    def _sum(self, parts):
        """Partial Gram matrices of all chunks -> one scaled matrix.  With several chunks (host-resident rows)
        every chunk but the last was already reduced to ONE matrix while the next chunk was on the link
        (add_chunk), so that what follows the last copy is one small sum, not a pass over all partials."""
        if not parts:
            return torch.zeros(self.n2, dtype=torch.float64, device=self.device)
        if len(parts) == 1:
            return self._finish(parts[0], self.scale)
        sums = [q if q.dim() == 1 else self._finish(q, 1.0) for q in parts]
        return self._finish(torch.stack(sums, 0), self.scale)

    def add_chunk(self, Xc: torch.Tensor, yc: torch.Tensor) -> None:
        if Xc.shape[0] == 0:
            return
        # the kernels read raw float64 storage through these pointers
        Xc, yc = _dev_f64(Xc, "X"), _dev_f64(yc, "y").contiguous()
        if Xc.device != yc.device or Xc.dim() != 2 or yc.dim() != 1 or Xc.shape[0] != yc.shape[0]:
            raise LsSpaCudaError("CholQR2.add_chunk: X must be (n, p), y (n,), on one device")
        if Xc.stride(1) != 1:
            Xc = Xc.contiguous()
        self.device = Xc.device
        self.chunks.append((Xc, yc))
        if self.parts1 and self.parts1[-1].dim() == 2:
            self.parts1[-1] = self._finish(self.parts1[-1], 1.0)     # the previous chunk, now that another follows
        self.parts1.append(self._rows(Xc, yc, None))

    def gram(self) -> torch.Tensor:
        """Scaled Gram matrix of this process's rows (zeros when it holds none)."""
        return self._sum(self.parts1)

    def factor(self, G: torch.Tensor, reg: float = 0.0, want_gram: bool = False):
        """want_gram: also leave the record of lsspa_lifts_gram (equilibrated Gram matrix, column scales,
        condition bound of the leading p x p block) in self.lift_gram -- the train side of a job."""
        lib = _lib()
        p, q = self.p, self.p + 1
        dev = G.device
        if reg != 0.0:
            check(lib.lsspa_gram_add_ridge(G.data_ptr(), p, float(reg), _stream()), "lsspa_gram_add_ridge")
            _count(1)
        self.R1 = torch.empty(q * q, dtype=torch.float64, device=dev)
        self.Rinv = torch.empty(int(lib.lsspa_gram_rinv_doubles(p)), dtype=torch.float64, device=dev)
        info = torch.zeros(2, dtype=torch.float64, device=dev)
        self.lift_gram = None
        if want_gram and lib.lsspa_lifts_chol_supported(p):
            self.lift_gram = torch.empty(int(lib.lsspa_lifts_gram_doubles(p)), dtype=torch.float64, device=dev)
        check(lib.lsspa_chol_factor_gram(G.data_ptr(), p, self.R1.data_ptr(), self.Rinv.data_ptr(), info.data_ptr(),
                                         _ptr(self.lift_gram), _stream()), "lsspa_chol_factor_gram")
        slot = torch.empty(tsqr_slot(p), dtype=torch.float64, device=dev)
        # one pass: slot = R1 (well conditioned: the second pass would only remove an orthogonality
        # defect of order eps * cond^2 <= 1e-10, and everything downstream depends on the factor
        # through R^T R = the Gram matrix, which Cholesky reproduces to eps)
        eye = torch.eye(q, dtype=torch.float64, device=dev)
        check(lib.lsspa_tri_product(eye.data_ptr(), self.R1.data_ptr(), p, G.data_ptr(), slot.data_ptr(),
                                    _stream()), "lsspa_tri_product")
        _count(2)
        self.G1 = G
        return slot, info

    def second_gram(self) -> torch.Tensor:
        """Gram matrix of Q1 = Z R1^-1 over this process's rows (pass 2 re-reads the resident chunks)."""
        return self._sum([self._rows(Xc, yc, self.Rinv) for Xc, yc in self.chunks])

    def finish_second(self, G2: torch.Tensor):
        lib = _lib()
        p = self.p
        dev = G2.device
        R2 = torch.empty_like(self.R1)
        Rinv2 = torch.empty_like(self.Rinv)
        info = torch.zeros(2, dtype=torch.float64, device=dev)
        check(lib.lsspa_chol_factor(G2.data_ptr(), p, R2.data_ptr(), Rinv2.data_ptr(), info.data_ptr(), _stream()),
              "lsspa_chol_factor")
        slot = torch.empty(tsqr_slot(p), dtype=torch.float64, device=dev)
        check(lib.lsspa_tri_product(R2.data_ptr(), self.R1.data_ptr(), p, self.G1.data_ptr(), slot.data_ptr(),
                                    _stream()), "lsspa_tri_product")
        _count(2)
        return slot, info

    def finish(self):
        """Single-process CholeskyQR2 without a ridge term -> (slot, info[2][2])."""
        slot, info1 = self.factor(self.gram(), 0.0)
        info = torch.zeros((2, 2), dtype=torch.float64, device=info1.device)
        info[0] = info1
        bad1, cond1 = (float(v) for v in info1.cpu())
        if bad1 == 0 and cond1 <= self.SINGLE_PASS_COND:
            info[1, 1] = 1.0
            return slot, info
        slot, info2 = self.finish_second(self.second_gram())
        info[1] = info2
        return slot, info


def cholqr2_factor(chunks, p: int, divisor: float):
    """One-shot helper: CholeskyQR2 over a list of device chunks [(X, y), ...]."""
    f = CholQR2(p, divisor)
    for Xc, yc in chunks:
        f.add_chunk(Xc, yc)
    return f.finish()


def gram_big_supported(p: int) -> bool:
    return bool(_lib().lsspa_gram_big_supported(p))


class GramBig:
    """One-pass Gram reduction of [X | y] / divisor for wide problems (p + 1 > 112; csrc/gram_big.cu):
    add_chunk() accumulates the dense (p+1)^2 Gram matrix of this process's rows, gram() hands it
    out (a multi-GPU job all-reduces it), factor(G, reg) runs the blocked Cholesky factorisation
    (csrc/lifts_big.cu with a batch of one) -> (slot, info) with info (device) = [bad pivot, nan]:
    the caller judges the conditioning from the factor itself (TrainSide)."""

    def __init__(self, p: int, divisor: float, device):
        self.p, self.scale, self.device = p, 1.0 / (float(divisor) ** 2), device
        self.G = torch.zeros((p + 1) * (p + 1), dtype=torch.float64, device=device)
        self.part_doubles = int(_lib().lsspa_gram_big_part_doubles(p))

    def add_chunk(self, Xc: torch.Tensor, yc: torch.Tensor) -> None:
        if Xc.shape[0] == 0:
            return
        Xc, yc = _dev_f64(Xc, "X"), _dev_f64(yc, "y").contiguous()
        if Xc.stride(1) != 1:
            Xc = Xc.contiguous()
        lib = _lib()
        n = Xc.shape[0]
        nsplit = lib.lsspa_gram_big_num_splits(self.p, n)
        parts = torch.empty((nsplit, self.part_doubles), dtype=torch.float64, device=Xc.device)
        check(lib.lsspa_gram_big_rows(Xc.data_ptr(), Xc.stride(0), yc.data_ptr(), n, self.p, parts.data_ptr(), nsplit,
                                      _stream()), "lsspa_gram_big_rows")
        check(lib.lsspa_gram_big_accumulate(parts.data_ptr(), nsplit, self.p, self.G.data_ptr(), _stream()),
              "lsspa_gram_big_accumulate")
        _count(2)

    def gram(self) -> torch.Tensor:
        return self.G

    def factor(self, G: torch.Tensor, reg: float = 0.0):
        lib = _lib()
        p = self.p
        nbytes = int(lib.lsspa_gram_big_factor_workspace_bytes(p))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=G.device)
        flag = torch.zeros(1, dtype=torch.int32, device=G.device)
        slot = torch.empty(tsqr_slot(p), dtype=torch.float64, device=G.device)
        check(lib.lsspa_gram_big_factor(G.data_ptr(), p, self.scale, float(reg), slot.data_ptr(), ws.data_ptr(), nbytes,
                                        flag.data_ptr(), _stream()), "lsspa_gram_big_factor")
        _count(2 + 2 * ((p + 64) // 64))
        info = torch.stack([flag.to(torch.float64)[0], torch.tensor(float("nan"), dtype=torch.float64, device=G.device)])
        return slot, info


def split_factor(slot: torch.Tensor, p: int):
    """(slot,) -> R (p,p) row-major upper triangular, c (p,), sum of squares of the y column."""
    q = p + 1
    T = slot[: q * q].view(q, q)
    return T[:p, :p], T[:p, p], slot[q * q]


def ridge_factor(p: int, reg: float, device) -> torch.Tensor:
    """The sqrt(reg) * I rows of the train block (reference ls_spa/ls_spa.py:310) as one more
    triangular factor in slot layout."""
    q = p + 1
    slot = torch.zeros(tsqr_slot(p), dtype=torch.float64, device=device)
    T = slot[: q * q].view(q, q)
    T.diagonal()[:p] = math.sqrt(reg)
    return slot


# ---------------------------------------------------------------------------
# permutation sources
# ---------------------------------------------------------------------------
def perms_exact(p: int, first_rank: int, count: int, device) -> torch.Tensor:
    out = torch.empty((count, p), dtype=torch.int32, device=device)
    check(_lib().lsspa_perms_exact(p, first_rank, count, out.data_ptr(), _stream()), "lsspa_perms_exact")
    _count(1)
    return out


def perms_pcg64(p: int, gen_state: torch.Tensor, count: int, status_flag: torch.Tensor) -> torch.Tensor:
    """gen_state: int64/uint64 CUDA tensor of 6 words, advanced in place."""
    lib = _lib()
    out = torch.empty((count, p), dtype=torch.int32, device=gen_state.device)
    if count == 0:
        return out
    nbytes = lib.lsspa_perms_pcg64_workspace_bytes(p, count)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=gen_state.device)
    check(lib.lsspa_perms_pcg64(p, gen_state.data_ptr(), count, out.data_ptr(), ws.data_ptr(), nbytes,
                                status_flag.data_ptr(), _stream()), "lsspa_perms_pcg64")
    _count(7)
    return out


def perms_sobol_argsort(p, sv, shift, bits, first_index, count) -> torch.Tensor:
    out = torch.empty((count, p), dtype=torch.int32, device=sv.device)
    check(_lib().lsspa_perms_sobol_argsort(p, sv.data_ptr(), shift.data_ptr(), bits, first_index, count,
                                           out.data_ptr(), _stream()), "lsspa_perms_sobol_argsort")
    _count(1)
    return out


def perms_permutohedron(p, sv, shift, bits, first_index, count) -> torch.Tensor:
    out = torch.empty((count, p), dtype=torch.int32, device=sv.device)
    check(_lib().lsspa_perms_permutohedron(p, sv.data_ptr(), shift.data_ptr(), bits, first_index, count,
                                           out.data_ptr(), _stream()), "lsspa_perms_permutohedron")
    _count(1)
    return out


def perms_validate(perms: torch.Tensor) -> None:
    """Raise ValueError unless every row of the int32 CUDA tensor `perms` is a bijection of {0..p-1}
    (the lift kernels index shared memory with these values).  One launch + one flag read."""
    if perms.numel() == 0:
        return
    perms = perms.contiguous()
    count, p = perms.shape
    flag = torch.zeros(1, dtype=torch.int32, device=perms.device)
    check(_lib().lsspa_perms_validate(p, perms.data_ptr(), count, flag.data_ptr(), _stream()), "lsspa_perms_validate")
    _count(1)
    bad = int(flag.item())
    if bad:
        raise ValueError(f"perms: row {bad - 1} is not a permutation of range({p})")


# ---------------------------------------------------------------------------
# per-permutation core
# ---------------------------------------------------------------------------
class TrainSide:
    """Train half of the reduced problem and what the Cholesky route derives from it: the
    equilibrated Gram matrix, its column scales and the condition bound that picks the route.
    It needs no test data, so a host-resident job builds it (and pre-factors permutations with it)
    while the test rows are still crossing PCIe."""

    def __init__(self, R_tr, c_tr, gram=None, cond=None):
        """gram / cond: the record lsspa_chol_factor_gram left behind and its condition bound, when the
        reduction has already produced them (no kernel, no host synchronisation here then)."""
        dev = R_tr.device
        self.p = p = int(R_tr.shape[0])
        # row-major (p,p) -> column-major storage == contiguous transpose
        self.R_tr_cm = _dev_f64(R_tr, "R_tr").t().contiguous()
        self.c_tr = _dev_f64(c_tr, "c_tr").contiguous()
        # Route of the per-permutation core: Cholesky of the permuted Gram matrix when the
        # train factor is well conditioned (error ~ eps * cond^2), Householder otherwise.
        self.gram = None
        self.scale = None
        self.cond_estimate = float("inf")
        self.use_chol = False
        forced = os.environ.get("LSSPA_LIFTS_IMPL", "")
        self.big = bool(_lib().lsspa_lifts_big_supported(p))     # wide problems: batched tile kernels (lifts_big.cu)
        if gram is not None and cond is not None and forced not in ("v1", "householder"):
            base = (p + 1) * (p + 1)
            self.gram = gram
            self.scale = gram[base + 8:base + 8 + p]
            self.set_cond(cond)
        elif forced not in ("v1", "householder") and (_lib().lsspa_lifts_chol_supported(p) or self.big):
            n = _lib().lsspa_lifts_gram_doubles(p)
            self.gram = torch.empty(n, dtype=torch.float64, device=dev)
            check(_lib().lsspa_lifts_gram(p, self.R_tr_cm.data_ptr(), self.c_tr.data_ptr(),
                                          self.gram.data_ptr(), _stream()), "lsspa_lifts_gram")
            _count(3)
            base = (p + 1) * (p + 1)
            info = self.gram[base:base + 2].cpu()
            self.cond_estimate = float(info[0])    # of the column-equilibrated train factor
            self.use_chol = forced == "chol" or self.cond_estimate <= CHOL_COND_LIMIT
            self.scale = self.gram[base + 8:base + 8 + p]


    def set_cond(self, cond) -> None:
        """Condition bound of the equilibrated train factor -> route.  A caller that builds this object
        before it has read the bound from the device (cond = inf) sets it here afterwards."""
        forced = os.environ.get("LSSPA_LIFTS_IMPL", "")
        if self.gram is None or forced in ("v1", "householder"):
            return
        self.cond_estimate = float(cond)
        self.use_chol = forced == "chol" or self.cond_estimate <= CHOL_COND_LIMIT


class ReducedProblem:
    """The p x p reduced factors in the layout the lift kernel wants (column-major)."""

    def __init__(self, R_tr, c_tr, R_te, c_te, y_norm_sq: float, train: TrainSide | None = None):
        train = train if train is not None else TrainSide(R_tr, c_tr)
        self.train = train
        dev = train.c_tr.device
        self.p = p = train.p
        R_te = _dev_f64(R_te, "R_te")
        if R_te.shape[0] < p:  # M < p: the reference keeps an M x p factor; zero rows change nothing
            pad = torch.zeros((p - R_te.shape[0], p), dtype=torch.float64, device=dev)
            R_te = torch.cat([R_te, pad], 0)
            c_te = torch.cat([_dev_f64(c_te, "c_te"), pad[:, 0]], 0)
        self.R_tr_cm, self.c_tr = train.R_tr_cm, train.c_tr
        self.R_te_cm = R_te.t().contiguous()
        self.c_te = _dev_f64(c_te, "c_te").contiguous()
        self.y_norm_sq = float(y_norm_sq)
        self._ws = None
        self.gram, self.cond_estimate, self.use_chol = train.gram, train.cond_estimate, train.use_chol
        if train.scale is not None:
            # the Cholesky route works on unit-norm train columns: scale the test columns alike
            self.R_te_scaled_cm = self.R_te_cm / train.scale.unsqueeze(1)

    def finalize(self, y_norm_sq: float) -> None:
        """For a problem assembled before the reduction's flags were read (engine.reduce_problem): the
        host-side scalars, once they are known."""
        self.y_norm_sq = float(y_norm_sq)
        self.cond_estimate, self.use_chol = self.train.cond_estimate, self.train.use_chol

    # device memory the tile workspace of the wide route may take (it is processed in passes)
    BIG_WS_BUDGET = 24 << 30

    def big_workspace(self, count: int, antithetical: bool):
        """(tile workspace, status flag) of the wide route, sized for as many samples as fit the budget."""
        dev = self.c_tr.device
        room = torch.cuda.get_device_properties(dev).total_memory - torch.cuda.memory_reserved(dev)
        held = 0 if getattr(self, "_big_ws", None) is None else self._big_ws.numel()
        budget = max(min(self.BIG_WS_BUDGET, (room + held) // 2), 1 << 20)
        nbytes = int(_lib().lsspa_lifts_big_workspace_bytes(self.p, count, 1 if antithetical else 0, budget))
        if getattr(self, "_big_ws", None) is None or self._big_ws.numel() < nbytes:
            self._big_ws = None
            self._big_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            self._big_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        return self._big_ws, self._big_flag

    def check(self) -> None:
        """Raise if a kernel of the wide route met a non-positive pivot (one flag read)."""
        flag = getattr(self, "_big_flag", None)
        if flag is not None and int(flag.item()) != 0:
            raise LsSpaCudaError("wide Cholesky route: non-positive pivot (the condition guard should have "
                                 "sent this problem to the Householder kernels)")

    def workspace(self, count: int):
        nbytes = _lib().lsspa_lifts_workspace_bytes(self.p, count)
        if nbytes == 0:
            return None, 0
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.c_tr.device)
        return self._ws, nbytes


def lifts(prob: ReducedProblem, perms: torch.Tensor, antithetical: bool, out: torch.Tensor | None = None):
    """perms (count, p) int32 CUDA -> lift rows (count, p) float64 (pair means if antithetical)."""
    if not perms.is_cuda or perms.dtype != torch.int32:
        raise LsSpaCudaError("perms must be an int32 CUDA tensor")
    perms = perms.contiguous()
    count, p = perms.shape
    if p != prob.p:
        raise LsSpaCudaError("perms width does not match the reduced problem")
    if out is None:
        out = torch.empty((count, p), dtype=torch.float64, device=perms.device)
    ws, nbytes = prob.workspace(count)
    trace = LIFT_TRACE
    if trace is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    global LIFT_ROUTE
    LIFT_ROUTE = "cholesky" if prob.use_chol else "householder"
    if prob.use_chol and prob.train.big:
        ws, flag = prob.big_workspace(count, antithetical)
        check(_lib().lsspa_lifts_big(p, prob.gram.data_ptr(), prob.R_te_scaled_cm.data_ptr(), prob.c_te.data_ptr(),
                                     prob.y_norm_sq, perms.data_ptr(), count, 1 if antithetical else 0,
                                     out.data_ptr(), ws.data_ptr(), ws.numel(), flag.data_ptr(), _stream()),
              "lsspa_lifts_big")
        _count(2 * (2 + 2 * ((p + 64) // 64)) - 1)     # per pass: gather, (diag + panel) per block row, cost
    elif prob.use_chol:
        check(_lib().lsspa_lifts_chol(p, prob.gram.data_ptr(), prob.R_te_scaled_cm.data_ptr(), prob.c_te.data_ptr(),
                                      prob.y_norm_sq, perms.data_ptr(), count, 1 if antithetical else 0,
                                      out.data_ptr(), _stream()), "lsspa_lifts_chol")
    else:
        check(_lib().lsspa_lifts(p, prob.R_tr_cm.data_ptr(), prob.c_tr.data_ptr(), prob.R_te_cm.data_ptr(),
                                 prob.c_te.data_ptr(), prob.y_norm_sq, perms.data_ptr(), count,
                                 1 if antithetical else 0, out.data_ptr(), _ptr(ws), nbytes, _stream()),
              "lsspa_lifts")
    _count(1)
    if trace is not None:
        e1.record()
        trace.append((e0, e1, count * (2 if antithetical else 1)))
    return out


def split_route_supported(p: int) -> bool:
    return int(_lib().lsspa_lifts_chol_factor_doubles(p)) > 0


def lifts_factor(train: TrainSide, perms: torch.Tensor, antithetical: bool) -> torch.Tensor:
    """First half of the Cholesky route (train side only): one factor block per permutation
    evaluation, (count * (2 if antithetical else 1), factor_doubles)."""
    perms = perms.contiguous()
    count, p = perms.shape
    fd = int(_lib().lsspa_lifts_chol_factor_doubles(p))
    evals = count * (2 if antithetical else 1)
    out = torch.empty((evals, fd), dtype=torch.float64, device=perms.device)
    check(_lib().lsspa_lifts_chol_factor(p, train.gram.data_ptr(), perms.data_ptr(), count,
                                         1 if antithetical else 0, out.data_ptr(), _stream()),
          "lsspa_lifts_chol_factor")
    _count(1)
    return out


def lifts_eliminate(prob: ReducedProblem, factors: torch.Tensor, perms: torch.Tensor, antithetical: bool,
                    out: torch.Tensor | None = None) -> torch.Tensor:
    """Second half: eliminate the test factor against the stored factors -> lift rows."""
    global LIFT_ROUTE
    LIFT_ROUTE = "cholesky"
    perms = perms.contiguous()
    count, p = perms.shape
    if out is None:
        out = torch.empty((count, p), dtype=torch.float64, device=perms.device)
    trace = LIFT_TRACE
    if trace is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(_lib().lsspa_lifts_chol_eliminate(p, factors.data_ptr(), prob.R_te_scaled_cm.data_ptr(),
                                            prob.c_te.data_ptr(), prob.y_norm_sq, perms.data_ptr(), count,
                                            1 if antithetical else 0, out.data_ptr(), _stream()),
          "lsspa_lifts_chol_eliminate")
    _count(1)
    if trace is not None:
        e1.record()
        trace.append((e0, e1, count * (2 if antithetical else 1)))
    return out


def theta_r2(prob: ReducedProblem):
    out = torch.empty(prob.p + 1, dtype=torch.float64, device=prob.c_tr.device)
    nbytes = _lib().lsspa_theta_r2_workspace_bytes(prob.p)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=prob.c_tr.device)
    check(_lib().lsspa_theta_r2(prob.p, prob.R_tr_cm.data_ptr(), prob.c_tr.data_ptr(), prob.R_te_cm.data_ptr(),
                                prob.c_te.data_ptr(), prob.y_norm_sq, out.data_ptr(), ws.data_ptr(), nbytes,
                                _stream()), "lsspa_theta_r2")
    _count(1)
    return out[: prob.p], out[prob.p]


# ---------------------------------------------------------------------------
# estimator
# ---------------------------------------------------------------------------
class Estimator:
    """Device-resident running statistics of the lift vectors: count, mean, biased covariance and
    the draw sums behind the error estimate.  The host side keeps the sample count, the error
    history and the stop decision (see engine.run_samples)."""

    def __init__(self, p: int, seed: int, estimate_errors: bool, device):
        lib = _lib()
        self.p = p
        self.estimate = bool(estimate_errors)
        self.seed = int(seed) & ((1 << 64) - 1)
        self.device = device
        self.state = torch.zeros(lib.lsspa_estimator_state_bytes(p) // 8, dtype=torch.float64, device=device)
        self.cur = 0              # which mean / G buffer is live
        self.count = 0            # samples folded in so far
        self.partial_doubles = int(lib.lsspa_estimator_partial_doubles(p))
        self.max_batches = int(lib.lsspa_estimator_max_batches(p))

    def partials(self, lift_rows: torch.Tensor, batch_desc) -> torch.Tensor:
        """batch_desc: list of (first_row, count, global_first_index) -> (nbatch, partial_doubles)."""
        nb = len(batch_desc)
        out = torch.empty((nb, self.partial_doubles), dtype=torch.float64, device=self.device)
        if nb == 0:
            return out
        desc = torch.tensor(batch_desc, dtype=torch.int64).reshape(nb, 3).to(self.device, non_blocking=True)
        check(_lib().lsspa_estimator_partials(self.p, lift_rows.data_ptr(), desc.data_ptr(), nb, self.seed,
                                              1 if self.estimate else 0, out.data_ptr(), _stream()),
              "lsspa_estimator_partials")
        _count(3 if self.estimate else 2)
        return out

    def snapshot(self):
        return self.state.clone(), self.cur, self.count

    # -- hierarchical merging (multi-GPU): a rank folds its own run of batches into a scratch
    #    estimator, ships the result as ONE partial block, and every rank folds the W run totals.
    def scratch(self) -> "Estimator":
        """A second estimator of the same shape (allocated once, reused)."""
        sc = getattr(self, "_scratch", None)
        if sc is None:
            sc = Estimator.__new__(Estimator)
            sc.__dict__.update({k: v for k, v in self.__dict__.items() if k != "_scratch"})
            sc.state = torch.zeros_like(self.state)
            sc.cur, sc.count = 0, 0
            self._scratch = sc
        return sc

    def reset(self) -> None:
        self.state.zero_()
        self.cur, self.count = 0, 0

    def copy_from(self, other: "Estimator") -> None:
        self.state.copy_(other.state)
        self.cur, self.count = other.cur, other.count

    def block_total(self, partials: torch.Tensor, nb: int) -> torch.Tensor:
        """The merge of the first nb partial blocks as ONE block (parallel sums; what a rank ships)."""
        out = torch.empty(self.partial_doubles, dtype=torch.float64, device=self.device)
        flat = partials.reshape(-1, self.partial_doubles)
        check(_lib().lsspa_estimator_block_total(self.p, _ptr(flat) if nb > 0 else 0, nb, 1 if self.estimate else 0,
                                                 out.data_ptr(), _stream()), "lsspa_estimator_block_total")
        _count(2)
        return out

    def export_block(self) -> torch.Tensor:
        """The whole state as one partial block {n, mean, M2 = n * biased cov, G, S} (the layout of
        lsspa_estimator_partials), so that it can be folded into another state by absorb()."""
        p, st, cur = self.p, self.state, self.cur
        o_g, o_cov, o_s = 2 * p, 2 * p + 2 * ERR_DRAWS, 2 * p + 2 * ERR_DRAWS + p * p
        blk = torch.zeros(self.partial_doubles, dtype=torch.float64, device=self.device)
        blk[0] = float(self.count)
        blk[8:8 + p] = st[cur * p:(cur + 1) * p]
        blk[8 + p:8 + p + p * p] = st[o_cov:o_s] * float(self.count)
        blk[8 + p + p * p:8 + p + p * p + ERR_DRAWS] = st[o_g + cur * ERR_DRAWS:o_g + (cur + 1) * ERR_DRAWS]
        blk[8 + p + p * p + ERR_DRAWS:] = st[o_s:]
        return blk

    def restore(self, snap) -> None:
        self.state.copy_(snap[0])
        self.cur, self.count = snap[1], snap[2]

    def absorb(self, partials: torch.Tensor, slots, counts, own=(0, 0), emit: bool = False,
               every_feature: bool = False):
        """Fold the batches whose partial blocks are partials[slots[b]] (counts[b] samples each), in
        order.  With emit=True returns (overall[own1-own0], per_feature[own1-own0, p]) device tensors
        holding the 0.95-quantile error estimates after each batch b in [own0, own1).  The sample loop
        only ever reads the per-feature errors of the LAST owned batch (the reference keeps
        attribution_errors of the batch it ends on, ls_spa/ls_spa.py:222-236), so by default only that
        row of per_feature is computed (the others are NaN) and the squared draws stay inside the fused
        kernel; every_feature=True takes the two-kernel route that fills every row."""
        lib = _lib()
        nb = len(slots)
        own0, own1 = own
        emit = emit and self.estimate and own1 > own0
        overall = feat = None
        if emit:
            overall = torch.empty(own1 - own0, dtype=torch.float64, device=self.device)
            feat = torch.full((own1 - own0, self.p), float("nan"), dtype=torch.float64, device=self.device)
        pos = 0
        flat = partials.reshape(-1, self.partial_doubles)
        while pos < nb:
            n = min(self.max_batches, nb - pos)
            smap = torch.tensor(slots[pos:pos + n], dtype=torch.int32).to(self.device, non_blocking=True)
            o0, o1 = max(own0, pos) - pos, min(own1, pos + n) - pos
            if emit and o1 > o0 and not every_feature:
                lo = pos + o0 - own0
                fb = o1 - 1 if pos + o1 == own1 else -1           # the chunk that holds the last owned batch
                ws = torch.empty(int(lib.lsspa_estimator_errors_workspace_doubles(self.p, o1 - o0)),
                                 dtype=torch.float64, device=self.device)
                check(lib.lsspa_estimator_absorb_errors(self.state.data_ptr(), self.p, self.cur, float(self.count),
                                                        flat.data_ptr(), smap.data_ptr(), n, o0, o1, fb,
                                                        overall[lo:].data_ptr(),
                                                        feat[own1 - own0 - 1:].data_ptr() if fb >= 0 else None,
                                                        ws.data_ptr(), _stream()), "lsspa_estimator_absorb_errors")
                _count(6)
            else:
                zsq = None
                if emit and o1 > o0:
                    zsq = torch.empty((o1 - o0, self.p + 1, ERR_DRAWS), dtype=torch.float64, device=self.device)
                check(lib.lsspa_estimator_absorb(self.state.data_ptr(), self.p, self.cur, float(self.count),
                                                 flat.data_ptr(), smap.data_ptr(), n, max(o0, 0), max(o1, 0),
                                                 _ptr(zsq), 1 if self.estimate else 0, _stream()),
                      "lsspa_estimator_absorb")
                _count(1)
                if zsq is not None:
                    lo = pos + o0 - own0
                    check(lib.lsspa_estimator_quantiles(self.p, zsq.data_ptr(), o1 - o0, overall[lo:].data_ptr(),
                                                        feat[lo:].data_ptr(), _stream()), "lsspa_estimator_quantiles")
                    _count(2)
            self.cur ^= 1
            self.count += int(sum(counts[pos:pos + n]))
            pos += n
        return overall, feat

    def read(self, want_cov: bool = False):
        """-> dict(count, mean[, cov]) as numpy (one device->host sync)."""
        p = self.p
        mean = self.state[self.cur * p:(self.cur + 1) * p].cpu().numpy().copy()
        res = dict(count=self.count, mean=mean)
        if want_cov:
            off = 2 * p + 2 * ERR_DRAWS
            res["cov"] = self.state[off:off + p * p].reshape(p, p).cpu().numpy().copy()
        return res


def error_draws_quantiles(cov: torch.Tensor, seed: int):
    """error_estimates (ls_spa/ls_spa.py:321-341) for an explicit covariance: (per-feature, overall)
    0.95 quantiles over 1024 device draws of N(0, cov)."""
    cov = _dev_f64(cov, "cov").contiguous()
    p = int(cov.shape[0])
    if cov.dim() != 2 or cov.shape[1] != p:
        raise LsSpaCudaError("cov must be square")
    zsq = torch.empty((1, p + 1, ERR_DRAWS), dtype=torch.float64, device=cov.device)
    ws = torch.empty(p * p, dtype=torch.float64, device=cov.device)
    lib = _lib()
    check(lib.lsspa_error_draws(p, cov.data_ptr(), int(seed) & ((1 << 64) - 1), zsq.data_ptr(), ws.data_ptr(),
                                _stream()), "lsspa_error_draws")
    overall = torch.empty(1, dtype=torch.float64, device=cov.device)
    feat = torch.empty((1, p), dtype=torch.float64, device=cov.device)
    check(lib.lsspa_estimator_quantiles(p, zsq.data_ptr(), 1, overall.data_ptr(), feat.data_ptr(), _stream()),
          "lsspa_estimator_quantiles")
    _count(4)
    return feat[0], overall[0]


def prefix_means(lift_rows: torch.Tensor, carry_sum: torch.Tensor, carry_count: int, out: torch.Tensor) -> None:
    rows, p = lift_rows.shape
    check(_lib().lsspa_prefix_means(p, lift_rows.data_ptr(), rows, carry_sum.data_ptr(), float(carry_count),
                                    out.data_ptr(), _stream()), "lsspa_prefix_means")
    _count(1)


def merge_moments(mean, cov, old_n, new_mean, new_cov, new_n) -> None:
    """In-place device Chan merge of (mean, biased cov); cov / new_cov may be None."""
    p = mean.numel()
    check(_lib().lsspa_merge_moments(p, mean.data_ptr(), _ptr(cov), float(old_n), new_mean.data_ptr(),
                                     _ptr(new_cov), float(new_n), _stream()), "lsspa_merge_moments")
    _count(2 if cov is not None else 1)
