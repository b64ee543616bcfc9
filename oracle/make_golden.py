"""Generate tests/golden/*.npz from the UNMODIFIED reference (run in the build
container only: /root/reference does not exist on the GPU box).

    PYTHONDONTWRITEBYTECODE=1 OPENBLAS_NUM_THREADS=1 python oracle/make_golden.py

Everything stored is an output of the reference's own functions
(``ls_spa.ls_spa``, ``square_shapley``, ``reduce_data``) or of the numpy/scipy
generators its drivers use, on seeded inputs that the tests can regenerate with
``oracle.samplers_oracle.gen_data``.  The synthetic inputs are rounded to float32
before the reference sees them and are stored as float32 (exactly representable), so
the tests never have to regenerate them: ``multivariate_normal(method="svd")`` goes through
LAPACK and is not reproducible across CPUs.
"""

from __future__ import annotations

import itertools
import os
import sys

import numpy as np

REF = "/root/reference"
sys.path.insert(0, REF)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ls_spa as ref_pkg  # noqa: E402  (the reference package)

ref_mod = sys.modules["ls_spa.ls_spa"]
from oracle import samplers_oracle as so  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def f32_exact(*arrays):
    """Round to float32 and return float64 arrays holding exactly those values."""
    return tuple(np.asarray(a, dtype=np.float32).astype(np.float64) for a in arrays)


def results_dict(res, prefix=""):
    d = {
        prefix + "attribution": res.attribution,
        prefix + "theta": res.theta,
        prefix + "overall_error": np.float64(res.overall_error),
        prefix + "attribution_errors": res.attribution_errors,
        prefix + "r_squared": np.float64(res.r_squared),
        prefix + "error_history": res.error_history,
    }
    if res.attribution_history is not None:
        d[prefix + "attribution_history"] = res.attribution_history
    return d


def toy():
    z = np.load(os.path.join(REF, "data", "toy_data.npz"))
    Xtr, Xte, ytr, yte = (z[k] for k in ("X_train", "X_test", "y_train", "y_test"))
    out = {"X_train": Xtr, "X_test": Xte, "y_train": ytr, "y_test": yte}
    out.update(results_dict(ref_pkg.ls_spa(Xtr, Xte, ytr, yte), "default_"))
    out.update(results_dict(ref_pkg.ls_spa(Xtr, Xte, ytr, yte, reg=0.1), "reg01_"))
    out.update(results_dict(ref_pkg.ls_spa(Xtr, Xte, ytr, yte, return_attribution_history=True), "hist_"))
    R_tr, R_te, c_tr, c_te = ref_mod.reduce_data(Xtr, Xte, ytr, yte, 0.0)
    ynsq = np.linalg.norm(yte) ** 2
    perms = np.array(list(itertools.permutations(range(3))))
    lifts = np.array([ref_mod.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm) for pm in perms])
    out.update(R_tr=R_tr, R_te=R_te, c_tr=c_tr, c_te=c_te, y_norm_sq=ynsq, perms=perms, lifts=lifts)
    out["default_repr"] = np.array(repr(ref_pkg.ls_spa(Xtr, Xte, ytr, yte)))
    wide = ref_pkg.ShapleyResults(np.arange(7) / 3.0, -np.arange(7) / 7.0, 1.25e-3, np.zeros(7), 0.5,
                                  np.zeros(0), None)
    out["wide_repr"] = np.array(repr(wide))
    np.savez_compressed(os.path.join(OUT, "toy.npz"), **out)
    print("toy attribution", out["default_attribution"])


def synthetic(tag, p, n, m, reg, k, seed_data=42, seed_perm=42, exact_small=False):
    rng = np.random.default_rng(seed_data)
    conditioning = 20.0 if p >= 20 else float(p)        # keep >=1 latent factor at small p
    Xtr, Xte, ytr, yte = f32_exact(*so.gen_data(rng, p, n, m, conditioning=conditioning)[:4])
    R_tr, R_te, c_tr, c_te = ref_mod.reduce_data(Xtr, Xte, ytr, yte, reg)
    ynsq = np.linalg.norm(yte) ** 2
    out = dict(p=p, n=n, m=m, reg=reg, conditioning=conditioning, seed_data=seed_data,
               X_train=Xtr.astype(np.float32), X_test=Xte.astype(np.float32),
               y_train=ytr.astype(np.float32), y_test=yte.astype(np.float32),
               R_tr=R_tr, R_te=R_te, c_tr=c_tr, c_te=c_te, y_norm_sq=ynsq)
    methods = {}
    methods["random"] = so.perms_random(p, k, seed_perm)
    methods["argsort"] = so.perms_argsort(p, k, seed_perm)[0]
    methods["permutohedron"] = so.perms_permutohedron(p, k, seed_perm)[0]
    if exact_small:
        methods["exact"] = so.perms_exact(p, k, first=0)
    for name, perms in methods.items():
        out[f"perms_{name}"] = perms.astype(np.int16)
        lifts = np.array([ref_mod.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm) for pm in perms])
        out[f"lifts_{name}"] = lifts
        for anti in (False, True):
            res = ref_pkg.ls_spa(Xtr, Xte, ytr, yte, reg=reg, perms=list(perms), tolerance=0.0,
                                 batch_size=max(k // 4, 2), antithetical=anti,
                                 return_attribution_history=True)
            out.update(results_dict(res, f"{name}_anti{int(anti)}_"))
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print(tag, "r2", out["random_anti0_r_squared"], "sum attr", out["random_anti0_attribution"].sum())


def exact_p7():
    """The reference's own exact path (p < 9): all 5040 permutations."""
    p, n, m = 7, 400, 300
    rng = np.random.default_rng(7)
    Xtr, Xte, ytr, yte = f32_exact(*so.gen_data(rng, p, n, m, conditioning=float(p))[:4])
    res = ref_pkg.ls_spa(Xtr, Xte, ytr, yte, reg=0.05, return_attribution_history=False)
    R_tr, R_te, c_tr, c_te = ref_mod.reduce_data(Xtr, Xte, ytr, yte, 0.05)
    out = dict(p=p, n=n, m=m, reg=0.05, conditioning=float(p), seed_data=7,
               X_train=Xtr.astype(np.float32), X_test=Xte.astype(np.float32),
               y_train=ytr.astype(np.float32), y_test=yte.astype(np.float32),
               R_tr=R_tr, R_te=R_te, c_tr=c_tr, c_te=c_te, y_norm_sq=np.linalg.norm(yte) ** 2)
    out.update(results_dict(res, "default_"))
    np.savez_compressed(os.path.join(OUT, "exact_p7.npz"), **out)
    print("exact_p7 attribution", res.attribution)


def streams():
    """Permutation streams alone, for the bit-exactness tests of the device generators."""
    out = {}
    for p, k in ((9, 64), (37, 64), (100, 512), (1000, 8)):
        out[f"random_p{p}_seed42"] = so.perms_random(p, k, 42).astype(np.int16)
    out["random_p100_seed7"] = so.perms_random(100, 128, 7).astype(np.int16)
    for p, k, seed in ((10, 256, 42), (100, 1024, 42), (100, 256, 7), (1000, 16, 42)):
        out[f"argsort_p{p}_seed{seed}"] = so.perms_argsort(p, k, seed)[0].astype(np.int16)
    for p, k, seed in ((10, 256, 42), (11, 64, 42), (100, 1024, 42), (100, 256, 7), (1000, 16, 42)):
        out[f"permutohedron_p{p}_seed{seed}"] = so.perms_permutohedron(p, k, seed)[0].astype(np.int16)
    out["exact_p5"] = so.perms_exact(5).astype(np.int16)
    out["exact_p10_first"] = so.perms_exact(10, 512).astype(np.int16)
    out["exact_p10_at_3000000"] = so.perms_exact(10, 512, first=3_000_000).astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "streams.npz"), **out)
    print("streams", {k: v.shape for k, v in out.items()})


from oracle.samplers_oracle import WIDTHS, width_problem  # noqa: E402


def widths():
    """square_shapley of the reference at every tile count of the Cholesky lift kernel (p = 17..152);
    only the lifts are stored, the inputs come from width_problem()."""
    out = {}
    for p in WIDTHS:
        R_tr, R_te, c_tr, c_te, ynsq, perms = width_problem(p)
        cond = np.linalg.cond(R_tr / np.linalg.norm(R_tr, axis=0))
        assert cond < 300, (p, cond)
        out[f"lifts_p{p}"] = np.array([ref_mod.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm) for pm in perms])
        out[f"cond_p{p}"] = np.float64(cond)
    np.savez_compressed(os.path.join(OUT, "widths.npz"), **out)
    print("widths", {k: float(v) for k, v in out.items() if k.startswith("cond")})


def ill_conditioned(tag="ill_p64", p=64, n=400, m=300, k=12, cond=1e5):
    """Train/test features with singular values spread over `cond` in random directions (column
    equilibration cannot repair that): the route the condition guard sends to the Householder kernels."""
    rng = np.random.default_rng(64)
    V = np.linalg.qr(rng.standard_normal((p, p)))[0]
    s = np.geomspace(1.0, 1.0 / cond, p)
    Xtr = np.linalg.qr(rng.standard_normal((n, p)))[0] * s @ V.T * np.sqrt(n)
    Xte = np.linalg.qr(rng.standard_normal((m, p)))[0] * s @ V.T * np.sqrt(m)
    theta = V @ (rng.standard_normal(p) / np.sqrt(s))
    ytr = Xtr @ theta + 0.05 * rng.standard_normal(n)
    yte = Xte @ theta + 0.05 * rng.standard_normal(m)
    Xtr, Xte, ytr, yte = f32_exact(Xtr, Xte, ytr, yte)
    R_tr, R_te, c_tr, c_te = ref_mod.reduce_data(Xtr, Xte, ytr, yte, 0.0)
    ynsq = np.linalg.norm(yte) ** 2
    out = dict(p=p, n=n, m=m, reg=0.0, X_train=Xtr.astype(np.float32), X_test=Xte.astype(np.float32),
               y_train=ytr.astype(np.float32), y_test=yte.astype(np.float32),
               R_tr=R_tr, R_te=R_te, c_tr=c_tr, c_te=c_te, y_norm_sq=ynsq,
               cond_equilibrated=np.linalg.cond(R_tr / np.linalg.norm(R_tr, axis=0)), cond=np.linalg.cond(R_tr))
    perms = so.perms_random(p, k, 5)
    out["perms_random"] = perms.astype(np.int16)
    out["lifts_random"] = np.array([ref_mod.square_shapley(R_tr, R_te, c_tr, c_te, ynsq, pm) for pm in perms])
    for anti in (False, True):
        res = ref_pkg.ls_spa(Xtr, Xte, ytr, yte, reg=0.0, perms=list(perms), tolerance=0.0, batch_size=4,
                             antithetical=anti, return_attribution_history=True)
        out.update(results_dict(res, f"random_anti{int(anti)}_"))
    np.savez_compressed(os.path.join(OUT, f"{tag}.npz"), **out)
    print(tag, "cond", out["cond"], "equilibrated", out["cond_equilibrated"], "r2", out["random_anti0_r_squared"],
          "max|lift|", np.abs(out["lifts_random"]).max())


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1:          # only the named fixtures, e.g. `make_golden.py widths ill_conditioned`
        for name in sys.argv[1:]:
            globals()[name]()
        sys.exit(0)
    widths()
    ill_conditioned()
    toy()
    exact_p7()
    synthetic("syn_p10", p=10, n=500, m=400, reg=0.0, k=48, exact_small=True)
    synthetic("syn_p33", p=33, n=600, m=500, reg=1e-2, k=32)
    synthetic("syn_p100", p=100, n=700, m=500, reg=0.0, k=32)
    synthetic("syn_p100_reg", p=100, n=600, m=60, reg=1e-2, k=16, seed_data=5, seed_perm=11)
    synthetic("syn_p160", p=160, n=500, m=300, reg=1e-3, k=8, seed_data=3, seed_perm=3)
    streams()
