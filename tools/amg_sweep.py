"""Robustness sweep of the default solver policy (AMG from 1 M unknowns) over element types and physics: every model is
solved with the default method and with Jacobi-CG; prints iterations-free summary (time, agreement, hierarchy)."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
torch.set_default_device("cuda")
import torchfem_b200 as T  # noqa: E402
from torchfem_b200.elements import linear_to_quadratic  # noqa: E402
from torchfem_b200.materials import (IsotropicConductivity2D, IsotropicConductivity3D, IsotropicElasticity3D,  # noqa: E402
                                     IsotropicElasticityPlaneStress)
from torchfem_b200.mesh import cube_hexa, cube_tetra, rect_quad, rect_tri  # noqa: E402


def timed(fn):
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, time.perf_counter() - t


def mech(model, nodes, axis_len):
    model.constraints[nodes[:, 0] == 0.0, :] = True
    right = (nodes[:, 0] - axis_len).abs() < 1e-12
    model.constraints[right, 0] = True
    model.displacements[right, 0] = 0.05
    return model


def heat(model, nodes, axis_len):
    model.constraints[nodes[:, 0] == 0.0, 0] = True
    right = (nodes[:, 0] - axis_len).abs() < 1e-12
    model.constraints[right, 0] = True
    model.temperatures[right, 0] = 100.0
    return model


def cases(scale):
    s = scale
    yield "Hexa1 solid", lambda: (lambda n, e: mech(T.Solid(n, e, IsotropicElasticity3D(1000.0, 0.3)), n, 1.0))(*cube_hexa(int(71 * s), int(71 * s), int(71 * s)))
    yield "Tetra1 solid", lambda: (lambda n, e: mech(T.Solid(n, e, IsotropicElasticity3D(1000.0, 0.3)), n, 1.0))(*cube_tetra(int(71 * s), int(71 * s), int(71 * s)))
    yield "Hexa2 solid", lambda: (lambda n, e: mech(T.Solid(n, e, IsotropicElasticity3D(1000.0, 0.3)), n, 1.0))(*linear_to_quadratic(*cube_hexa(int(46 * s), int(46 * s), int(46 * s))))
    yield "Tetra2 solid", lambda: (lambda n, e: mech(T.Solid(n, e, IsotropicElasticity3D(1000.0, 0.3)), n, 1.0))(*linear_to_quadratic(*cube_tetra(int(36 * s), int(36 * s), int(36 * s))))
    yield "Quad1 planar", lambda: (lambda n, e: mech(T.Planar(n, e, IsotropicElasticityPlaneStress(1000.0, 0.3)), n, 2.0))(*rect_quad(int(1001 * s), int(501 * s), 2.0, 1.0))
    yield "Tria2 planar", lambda: (lambda n, e: mech(T.Planar(n, e, IsotropicElasticityPlaneStress(1000.0, 0.3)), n, 2.0))(*linear_to_quadratic(*rect_tri(int(501 * s), int(251 * s), 2.0, 1.0)))
    yield "Hexa1 heat", lambda: (lambda n, e: heat(T.SolidHeat(n, e, IsotropicConductivity3D(10.0)), n, 1.0))(*cube_hexa(int(101 * s), int(101 * s), int(101 * s)))
    yield "Quad2 heat", lambda: (lambda n, e: heat(T.PlanarHeat(n, e, IsotropicConductivity2D(10.0)), n, 2.0))(*linear_to_quadratic(*rect_quad(int(1001 * s), int(501 * s), 2.0, 1.0)))


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    for name, build in cases(scale):
        rec = {"case": name}
        try:
            model = build()
            rec["n_dofs"] = model.n_dofs
            rec["default_method"] = T.sparse.resolve_method(model.n_dofs, "cuda", None)
            out = {}
            for method in ("amgx", "cg"):
                for rep in range(2):
                    (u, *_), t = timed(lambda: model.solve(method=method, stol=1e-9))
                out[method] = u
                rec[f"{method}_s"] = round(t, 4)
            rec["rel_diff"] = float((out["amgx"] - out["cg"]).norm() / out["cg"].norm())
            M = model.pattern.sell_structure.amg_cache
            rec["levels"] = [lv.n for lv in M.levels]
            rec["agg_distance"] = [getattr(lv, "agg_distance", None) for lv in M.levels[:-1]]
            rec["operator_complexity"] = round(M.operator_complexity, 3)
            x, st = M.solve(torch.ones(model.n_dofs), rtol=1e-8, maxiter=500)
            rec["amg_iterations_rhs_ones"] = st["iterations"]
            del model, M, out
        except Exception as e:  # noqa: BLE001
            rec["error"] = f"{type(e).__name__}: {e}"[:300]
        torch.cuda.empty_cache()
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
