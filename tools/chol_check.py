"""Cholesky route of the per-permutation core against the Householder route and the goldens.

Prints, per problem: the condition estimate, the scaled difference of the two routes, the scaled
difference of each route to the reference's lifts (goldens only) and the timings.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tools")
from quick_bench import ev_time, synth_problem  # noqa: E402
from ls_spa_b200 import ops, samplers  # noqa: E402


def scaled(a, b):
    return float((a - b).abs().max() / b.abs().max())


def run(prob, perms, anti):
    outs = {}
    for name, flag in (("householder", False), ("chol", True)):
        prob.use_chol = flag
        buf = torch.empty((perms.shape[0], prob.p), dtype=torch.float64, device=perms.device)
        ops.lifts(prob, perms, anti, out=buf)
        torch.cuda.synchronize()
        outs[name] = buf
    return outs


def main():
    dev = torch.device("cuda")
    for name in ("syn_p100", "syn_p100_reg"):
        g = np.load(f"tests/golden/{name}.npz")
        prob = ops.ReducedProblem(*(torch.from_numpy(g[k]).to(dev) for k in ("R_tr", "c_tr", "R_te", "c_te")),
                                  float(np.sum(g["y_test"].astype(np.float64) ** 2)))
        for method in ("random", "argsort", "permutohedron"):
            if f"perms_{method}" not in g.files:
                continue
            perms = torch.from_numpy(g[f"perms_{method}"].astype(np.int32)).to(dev)
            ref = torch.from_numpy(g[f"lifts_{method}"]).to(dev)
            o = run(prob, perms, False)
            print(name, method, "cond", f"{prob.cond_estimate:.3g}",
                  "chol-vs-hh", f"{scaled(o['chol'], o['householder']):.2e}",
                  "hh-vs-ref", f"{scaled(o['householder'], ref):.2e}",
                  "chol-vs-ref", f"{scaled(o['chol'], ref):.2e}", flush=True)
    for p in (49, 56, 64, 71, 96, 100, 104, 117, 120, 128):
        prob = synth_problem(p, dev)
        perms = samplers.ArgsortSource(p, 7, None, dev).take(2048)
        for anti in (False, True):
            o = run(prob, perms, anti)
            print("synth p", p, "anti", anti, "cond", f"{prob.cond_estimate:.3g}",
                  "chol-vs-hh", f"{scaled(o['chol'], o['householder']):.2e}",
                  "nan", bool(torch.isnan(o['chol']).any()), flush=True)
    prob = synth_problem(100, dev)
    perms = samplers.PermutohedronSource(100, 42, None, dev).take(1 << 14)
    buf = torch.empty((perms.shape[0], 100), dtype=torch.float64, device=dev)
    for name, flag in (("householder", False), ("chol", True)):
        prob.use_chol = flag
        t = ev_time(lambda: ops.lifts(prob, perms, True, out=buf))
        n = 2 * perms.shape[0]
        print(name, f"{t:.3f} ms", f"{n / t * 1e3 / 1e6:.3f} M evals/s",
              f"{7 / 3 * 1e6 * n / t / 1e9:.2f} TF/s (7/3 p^3)", flush=True)


if __name__ == "__main__":
    main()
