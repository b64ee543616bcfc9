"""Permutation sources: host-side set-up, device-side generation.

Each source hands out consecutive chunks of the same pure stream the reference's
drivers would produce:

  exact          itertools.permutations(range(p))             ls_spa/ls_spa.py:171
  random         default_rng(seed).permutation(p) repeated     ls_spa/ls_spa.py:168,175
  argsort        argsort(Sobol(p, seed).random(1))             experiments/ground_truth_medium.py:70-71
  permutohedron  permutohedron_samples(MultivariateNormalQMC)  experiments/ground_truth_medium.py:56-67
  explicit       anything iterable handed in through perms=    ls_spa/ls_spa.py:176-177

The host only derives the generator constants from the seed with numpy / scipy
(PCG64 state of ``default_rng(seed)``; scipy's scrambled Sobol' direction numbers and
digital shift); the permutations themselves are produced by the CUDA kernels in
csrc/perms.cu, bit-exact with the host streams.
"""

from __future__ import annotations

import itertools
import math
import warnings

import numpy as np
import torch

from . import ops
from ._cabi import LsSpaCudaError

_MASK64 = (1 << 64) - 1


def _as_i64(words):
    """uint64 words -> int64 tensor with the same bit patterns."""
    return torch.from_numpy(np.array(words, dtype=np.uint64).view(np.int64))


class PermutationSource:
    """total: number of permutations the stream holds (None = unknown / unbounded)."""

    method = "?"
    random_access = False

    def __init__(self, p: int, total):
        self.p, self.total, self.position = p, total, 0

    def take(self, count: int) -> torch.Tensor:
        """Next `count` permutations as an int32 CUDA tensor (may be shorter at the end)."""
        raise NotImplementedError

    def skip(self, count: int) -> None:
        """Advance without materialising (random-access sources only)."""
        if not self.random_access:
            raise LsSpaCudaError(f"{self.method} source cannot skip")
        self.position += count

    def _clip(self, count):
        if self.total is not None:
            count = min(count, self.total - self.position)
        return max(int(count), 0)


class ExactSource(PermutationSource):
    method = "exact"
    random_access = True

    def __init__(self, p, device):
        if p > 20:
            raise LsSpaCudaError("exact enumeration is limited to p <= 20 (20! < 2^64)")
        super().__init__(p, math.factorial(p))
        self.device = device

    def take(self, count):
        count = self._clip(count)
        out = ops.perms_exact(self.p, self.position, count, self.device)
        self.position += count
        return out


class RandomSource(PermutationSource):
    method = "random"

    def __init__(self, p, seed, total, device):
        """seed: an int (the stream of ``default_rng(seed)``, reference :168) or a numpy
        ``Generator``: the device stream then continues exactly where that generator stands (the
        reference's experiment scripts draw data and permutations from one generator); call
        ``sync_generator`` afterwards to move the host generator past what the device consumed."""
        super().__init__(p, total)
        self.host_generator = seed if isinstance(seed, np.random.Generator) else None
        st = (self.host_generator if self.host_generator is not None
              else np.random.default_rng(seed)).bit_generator.state
        if st["bit_generator"] != "PCG64":
            raise LsSpaCudaError("the permutation stream needs a PCG64 bit generator (numpy's default_rng)")
        s, inc = st["state"]["state"], st["state"]["inc"]
        words = [s >> 64, s & _MASK64, inc >> 64, inc & _MASK64, int(st["has_uint32"]), int(st["uinteger"])]
        self.gen_state = _as_i64(words).to(device)
        self.flag = torch.zeros(1, dtype=torch.int32, device=device)

    def take(self, count):
        count = self._clip(count)
        out = ops.perms_pcg64(self.p, self.gen_state, count, self.flag)
        self.position += count
        return out

    def check(self):
        if int(self.flag.item()) != 0:
            raise LsSpaCudaError("PCG64 raw-draw budget exhausted on the device (should not happen)")

    def sync_generator(self, rng=None):
        """Write the device stream position back into the numpy generator it was started from."""
        rng = rng if rng is not None else self.host_generator
        if rng is None:
            return
        w = [int(v) & _MASK64 for v in self.gen_state.cpu().tolist()]
        st = rng.bit_generator.state
        st["state"]["state"] = (w[0] << 64) | w[1]
        st["has_uint32"], st["uinteger"] = int(w[4] != 0), int(w[5])
        rng.bit_generator.state = st


_SOBOL_CONST = {}


def sobol_tables(d: int, seed, spawn: bool = False, bits: int = 30):
    """scipy's scrambled Sobol' direction numbers and digital shift (uint32 (d, bits), (d,)) for an
    integer seed, built the way ``scipy.stats.qmc.Sobol(d, seed=seed)._scramble`` builds them
    (scipy/stats/_qmc.py:1812-1828 + ``_cscramble``: the same ``default_rng(seed)`` draws, unit-diagonal
    lower-triangular bit matrices, parity of row & direction number) but vectorised: ~1 ms at d = 100
    where constructing the scipy engine takes ~8 ms -- which used to leave the GPU idle between the
    reduction and the first lifts.  spawn=True: the stream ``MultivariateNormalQMC`` hands its engine
    (it passes a Generator, which QMCEngine spawns from, _qmc.py ``_initialize``).  Checked against
    the scipy engines by tests/test_sampler_models.py."""
    from scipy.stats import _sobol
    rng = np.random.default_rng(seed)
    if spawn:
        rng = rng.spawn(1)[0]
    sv = np.zeros((d, bits), dtype=np.uint64)
    _sobol._initialize_v(sv, dim=d, bits=bits)
    c = _SOBOL_CONST.get(bits)
    if c is None:
        w = np.uint64(1) << np.arange(bits, dtype=np.uint64)
        wr = np.ascontiguousarray(w[::-1])
        c = _SOBOL_CONST[bits] = (w, wr, np.tril(np.ones((bits, bits), dtype=np.uint64), -1),
                                  np.eye(bits, dtype=np.uint64), wr.astype(np.uint32))
    w, wr, tri, eye, wr32 = c
    shift = rng.integers(2, size=(d, bits), dtype=np.uint64) @ w
    ltm = rng.integers(2, size=(d, bits, bits), dtype=np.uint64)
    ltm *= tri
    ltm += eye                                           # _cscramble sets the diagonals to 1
    rows = (ltm @ wr).astype(np.uint32)                  # row p as an integer, entry k weighs 2^(bits-1-k)
    x = rows[:, :, None] & sv.astype(np.uint32)[:, None, :]
    par = (np.bitwise_count(x) & 1).astype(np.uint32)
    return (par * wr32[None, :, None]).sum(axis=1, dtype=np.uint32), shift.astype(np.uint32)


class _SobolBacked(PermutationSource):
    random_access = True

    def _upload_tables(self, d, seed, spawn, device):
        """Fast path for integer seeds and the 30-bit engines the reference's drivers use."""
        sv, shift = sobol_tables(d, int(seed), spawn=spawn, bits=30)
        self.bits = 30
        self.sv = torch.from_numpy(sv.view(np.int32)).to(device)
        self.shift = torch.from_numpy(shift.view(np.int32)).to(device)

    def _upload(self, engine, device):
        if engine.bits > 32:
            raise LsSpaCudaError("Sobol engines with more than 32 bits are not supported")
        self.bits = int(engine.bits)
        self.sv = torch.from_numpy(np.ascontiguousarray(engine._sv).astype(np.uint32).view(np.int32)).to(device)
        self.shift = torch.from_numpy(np.ascontiguousarray(engine._shift).astype(np.uint32).view(np.int32)).to(device)


class ArgsortSource(_SobolBacked):
    method = "argsort"

    def __init__(self, p, seed, total, device):
        super().__init__(p, total if total is not None else 2 ** 30)
        if isinstance(seed, (int, np.integer)) and hasattr(np, "bitwise_count"):
            self._upload_tables(p, seed, False, device)
        else:
            from scipy.stats.qmc import Sobol
            self._upload(Sobol(p, seed=seed), device)
        self.total = min(self.total, 2 ** self.bits)

    def take(self, count):
        count = self._clip(count)
        out = ops.perms_sobol_argsort(self.p, self.sv, self.shift, self.bits, self.position, count)
        self.position += count
        return out


class PermutohedronSource(_SobolBacked):
    method = "permutohedron"

    def __init__(self, p, seed, total, device):
        from scipy.stats.qmc import MultivariateNormalQMC
        if p < 2:
            raise LsSpaCudaError("permutohedron sampling needs p >= 2")
        super().__init__(p, total if total is not None else 2 ** 30)
        if isinstance(seed, (int, np.integer)) and hasattr(np, "bitwise_count"):
            self._upload_tables(2 * math.ceil((p - 1) / 2), seed, True, device)     # engine dimension: even
        else:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                qmc = MultivariateNormalQMC(np.zeros(p - 1), seed=seed, inv_transform=False)
            self._upload(qmc.engine, device)
        self.total = min(self.total, 2 ** self.bits)

    def take(self, count):
        count = self._clip(count)
        out = ops.perms_permutohedron(self.p, self.sv, self.shift, self.bits, self.position, count)
        self.position += count
        return out


class ExplicitSource(PermutationSource):
    """Permutations supplied by the caller (array, list, generator, CUDA tensor)."""

    method = "explicit"

    def __init__(self, p, perms, device):
        self.device = device
        self._tensor = None
        self._iter = None
        total = None
        if isinstance(perms, torch.Tensor):
            t = perms.reshape(-1, p)
            self._tensor = t.to(device=device, dtype=torch.int32)
            total = t.shape[0]
        elif isinstance(perms, np.ndarray) and perms.ndim == 2:
            if perms.size and (perms.min() < 0 or perms.max() >= p):
                raise ValueError(f"perms: entries must lie in range({p})")
            self._tensor = torch.from_numpy(np.ascontiguousarray(perms).astype(np.int32)).to(device)
            total = perms.shape[0]
        else:
            self._iter = iter(perms)
            try:
                total = len(perms)
            except TypeError:
                total = None
        super().__init__(p, total)
        self.exhausted = False

    @staticmethod
    def _validate(out):
        # the product path always runs on a CUDA device (ops.require_cuda); CPU tensors only occur in the
        # host-logic unit tests, whose oracle-backed stand-in indexes with numpy (IndexError on bad input)
        if out.is_cuda:
            ops.perms_validate(out)

    def take(self, count):
        """Every chunk handed out was checked on the device to consist of bijections of range(p):
        the reference leaves that to numpy's IndexError (ls_spa/ls_spa.py:165-167), here a bad
        index would address shared memory out of bounds."""
        if self._tensor is not None:
            count = self._clip(count)
            out = self._tensor[self.position:self.position + count]
            self.position += count
            self._validate(out)
            return out
        rows = list(itertools.islice(self._iter, count))
        if len(rows) < count:
            self.exhausted = True
            if self.total is None:
                self.total = self.position + len(rows)
        if not rows:
            return torch.empty((0, self.p), dtype=torch.int32, device=self.device)
        arr = np.asarray(rows).reshape(len(rows), -1)
        if arr.shape[1] != self.p:
            raise LsSpaCudaError("explicit permutations must have p entries each")
        self.position += len(rows)
        if arr.size and (arr.min() < 0 or arr.max() >= self.p):      # before the cast to int32 can wrap
            raise ValueError(f"perms: entries must lie in range({self.p})")
        out = torch.from_numpy(arr.astype(np.int32)).to(self.device)
        self._validate(out)
        return out


def make_source(method, p, seed, total, device, perms=None) -> PermutationSource:
    if perms is not None:
        return ExplicitSource(p, perms, device)
    if method == "exact":
        return ExactSource(p, device)
    if method == "random":
        return RandomSource(p, seed, total, device)
    if method == "argsort":
        return ArgsortSource(p, seed, total, device)
    if method == "permutohedron":
        return PermutohedronSource(p, seed, total, device)
    raise ValueError(f"unknown method {method!r}; expected 'random', 'permutohedron', 'argsort' or 'exact'")
