"""AMG-PCG on the benchmark cube: setup / solve timings, iteration counts and the hierarchy, next to Jacobi-PCG.
    python tools/amg_check.py --edge 64 100 150
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, nargs="+", default=[64])
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--jacobi", action="store_true")
    ap.add_argument("--hostprof", action="store_true", help="wall time per C-ABI / torch call of an un-synchronised setup")
    a = ap.parse_args()
    torch.set_default_dtype(torch.float64)
    import torchfem_b200 as T
    from torchfem_b200 import csr
    from torchfem_b200.amg import AMGPreconditioner
    from torchfem_b200.elements import Hexa1
    from torchfem_b200.materials import IsotropicElasticity3D

    dev = torch.device("cuda", 0)
    for E in a.edge:
        nodes, elements, con, disp = bench.build_problem(T, torch, E, dev)
        bref, w = bench.element_tables(Hexa1)
        C = IsotropicElasticity3D(torch.full((len(elements),), 1000.0, device=dev), torch.full((len(elements),), 0.3, device=dev)).C
        nodes, elements = nodes.to(dev), elements.to(dev)
        is_con = con.ravel().to(torch.uint8).to(dev)
        ubc = (disp.ravel() * con.ravel()).to(dev)
        p = csr.Pattern(elements, nodes.shape[0], 3)
        k = csr.integrate_k(T._lib.KIND_MECH, bref, w, nodes, elements, C)
        lift = torch.empty(p.n_dofs, device=dev)
        vals = csr.assemble(p, k, is_con, ubc=ubc, lift=lift)
        del k
        A = p.matrix(vals)
        b = lift
        A.sell()
        out = {"edge": E, "n_dofs": p.n_dofs, "nnz": p.nnz}
        for rep in range(2):
            amg, t_setup = timed(lambda: AMGPreconditioner(A))
            (x, st), t_solve = timed(lambda: amg.solve(b, rtol=a.rtol))
        os.environ["TFEM_AMG_TIMING"] = "1"
        amg_t = AMGPreconditioner(A)
        out["setup_phases_ms"] = {k: round(v, 2) for k, v in amg_t._timing.items()}
        amg_t._setup(A, symbolic=False)
        out["resetup_phases_ms"] = {k: round(v, 2) for k, v in amg_t._timing.items()}
        del amg_t
        os.environ.pop("TFEM_AMG_TIMING")
        _, t_resetup = timed(lambda: amg._setup(A, symbolic=False))
        (x, st), t_solve2 = timed(lambda: amg.solve(b, rtol=a.rtol))
        res = float((A.matvec(x) - b).norm() / b.norm())
        out.update({"amg_setup_ms": t_setup, "amg_resetup_ms": t_resetup, "amg_solve_ms": t_solve2,
                    "amg_iterations": st["iterations"], "amg_ms_per_iteration": t_solve2 / max(st["iterations"], 1),
                    "true_rel_residual": res, "levels": [(lv.n, lv.op.nblk * lv.d ** 2, "bcsr" if lv.op.use_bcsr else "sell") for lv in amg.levels],
                    "PR": [(lv.P.nblk, "bcsr" if lv.P.use_bcsr else "sell", "bcsr" if lv.R.use_bcsr else "sell")
                           for lv in amg.levels[:-1]],
                    "operator_complexity": amg.operator_complexity,
                    "mis_rounds": [getattr(lv, "mis_rounds", None) for lv in amg.levels[:-1]],
                    "rho": [getattr(lv, "rho", None) for lv in amg.levels[:-1]],
                    "launches": st["launches"],
                    "amg_dofs_per_s_setup_plus_solve": p.n_dofs / ((t_setup + t_solve2) * 1e-3)})
        if a.hostprof:
            import collections
            from torchfem_b200 import _lib as LL, amg as amg_mod

            acc = collections.defaultdict(float)

            class Wrap:
                def __init__(self, lib):
                    self._lib = lib

                def __getattr__(self, name):
                    fn = getattr(self._lib, name)

                    def call(*args):
                        t0 = time.perf_counter()
                        r = fn(*args)
                        acc[name] += (time.perf_counter() - t0) * 1e3
                        return r
                    return call

            real_lib, real_empty, real_inv, real_zeros = LL.lib, torch.empty, torch.linalg.inv, torch.zeros

            def t_wrap(fn, name):
                def call(*args, **kw):
                    t0 = time.perf_counter()
                    r = fn(*args, **kw)
                    acc[name] += (time.perf_counter() - t0) * 1e3
                    return r
                return call
            LL.lib = Wrap(real_lib)
            torch.empty, torch.zeros = t_wrap(real_empty, "torch.empty"), t_wrap(real_zeros, "torch.zeros")
            torch.linalg.inv = t_wrap(real_inv, "torch.linalg.inv")
            try:
                _, t_un = timed(lambda: AMGPreconditioner(A))
            finally:
                LL.lib, torch.empty, torch.zeros, torch.linalg.inv = real_lib, real_empty, real_zeros, real_inv
            out["hostprof_total_ms"] = t_un
            out["hostprof_ms"] = {k: round(v, 2) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:14]}
        if a.jacobi:
            for rep in range(2):
                (xj, _, sj), t_j = timed(lambda: csr.krylov_solve(A, b, method="cg", rtol=a.rtol))
            out.update({"jacobi_solve_ms": t_j, "jacobi_iterations": sj["iterations"],
                        "rel_diff_amg_vs_jacobi": float((x - xj).norm() / xj.norm())})
        print(json.dumps(out), flush=True)
        del amg, A, p, vals, x
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
