"""`Assembly.solve` on the GPU at benchmark size: two N^3-node Hexa1 blocks tied at a face (and, second case, a
reference point driving the top face), next to the monolithic bar solved by `Solid.solve` with the same method.
    python tools/assembly_check.py --nodes 61 --method cg
Prints one JSON line per case: wall times (first call = elimination map + symbolic SpGEMM phases + solve; second call
= numeric phases + solve), the phases of the reduction on their own, and the agreement with the monolithic solution.
"""
import argparse
import json
import os
import sys
import time
import traceback

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    return out, round((time.perf_counter() - t0) * 1e3, 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=61)
    ap.add_argument("--method", default="cg")
    ap.add_argument("--stol", type=float, default=1e-8)
    ap.add_argument("--cases", nargs="+", default=["tie", "point"])
    ap.add_argument("--scalar", action="store_true", help="scalar (d = 1) operators instead of node blocks")
    a = ap.parse_args()
    torch.set_default_dtype(torch.float64)
    torch.set_default_device("cuda")
    import torchfem_b200 as T
    from torchfem_b200.assembly import EMPTY, _Elimination
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    N, method, stol = a.nodes, a.method, a.stol
    mat = IsotropicElasticity3D(1000.0, 0.3)
    load = 1.0 / N ** 2

    nodes, elements = cube_hexa(N, N, 2 * N - 1, 1.0, 1.0, 2.0)
    mono = T.Solid(nodes, elements, mat)
    mono.constraints[nodes[:, 2] == 0.0] = True
    mono.forces[nodes[:, 2] == 2.0, 2] = load
    mono.forces[nodes[:, 2] == 2.0, 0] = 0.2 * load
    _, t1 = timed(lambda: mono.solve(method=method, stol=stol))
    (u_ref, *_), t2 = timed(lambda: mono.solve(method=method, stol=stol))
    print(json.dumps({"case": "monolithic", "n_dofs": mono.n_dofs, "method": method, "stol": stol,
                      "solve_ms_first": t1, "solve_ms_warm": t2}), flush=True)

    n_a, e_a = cube_hexa(N, N, N, 1.0, 1.0, 1.0)
    n_b = n_a + torch.tensor([0.0, 0.0, 1.0])
    face_a, face_b = torch.isclose(n_a[:, 2], torch.tensor(1.0)), torch.isclose(n_b[:, 2], torch.tensor(1.0))
    top_b = n_b[:, 2] == 2.0

    def parts():
        pa, pb = T.Solid(n_a, e_a, mat), T.Solid(n_b, e_a.clone(), mat)
        pa.constraints[n_a[:, 2] == 0.0] = True
        return pa, pb

    try:
        if "tie" not in a.cases:
            raise KeyError
        pa, pb = parts()
        pb.forces[top_b, 2] = load
        pb.forces[top_b, 0] = 0.2 * load
        asm = T.Assembly([pa, pb])
        asm.node_blocks = not a.scalar
        _, tc = timed(lambda: asm.coupling(pb, face_b, pa, face_a))
        _, t1 = timed(lambda: asm.solve(method=method, stol=stol))
        (u, *_), t2 = timed(lambda: asm.solve(method=method, stol=stol))
        scale = float(u_ref.abs().max())
        lower, upper = nodes[:, 2] <= 1.0 + 1e-12, nodes[:, 2] >= 1.0 - 1e-12
        err = max(float((u[0] - u_ref[lower]).abs().max()), float((u[1] - u_ref[upper]).abs().max())) / scale
        # the phases of the reduction on their own
        elim, t_map = timed(lambda: _Elimination(asm, asm.node_blocks))
        blocks = [p.assemble_matrix(p.k0(), EMPTY) for p in asm.parts]
        con = torch.nonzero(torch.cat([p.constraints.ravel() for p in asm.parts])[elim.retained]).ravel()
        K, t_first = timed(lambda: elim.reduced(blocks, con))
        _, t_numeric = timed(lambda: elim.reduced([b * 1.0 for b in blocks], con))
        print(json.dumps({"case": "tie", "block_size": elim.d, "n_dofs": asm.n_dofs, "n_retained": elim.n_retained, "nnz_reduced": K.nnz,
                          "coupling_ms": tc, "solve_ms_first": t1, "solve_ms_warm": t2, "rel_err_vs_monolithic": err,
                          "elimination_map_ms": t_map, "reduce_symbolic_plus_numeric_ms": t_first,
                          "reduce_numeric_ms": t_numeric}), flush=True)
    except KeyError:
        pass
    except Exception:
        traceback.print_exc()

    try:
        if "point" not in a.cases:
            raise KeyError
        pa, pb = parts()
        point = T.ReferencePoint([0.5, 0.5, 2.5])
        point.forces[0, 2], point.forces[0, 0], point.forces[0, 4] = 1.0, 0.2, 0.05
        asm = T.Assembly([pa, pb, point])
        asm.node_blocks = not a.scalar
        asm.coupling(pb, face_b, pa, face_a)
        asm.coupling(pb, top_b, point)
        _, t1 = timed(lambda: asm.solve(method=method, stol=stol))
        (u, f, *_), t2 = timed(lambda: asm.solve(method=method, stol=stol))
        u_p, theta = u[2][0, :3], u[2][0, 3:]
        rigid = u_p + torch.cross(theta.expand(int(top_b.sum()), 3), n_b[top_b] - point.nodes[0], dim=-1)
        elim = asm._elimination
        K, t_numeric = timed(lambda: elim.reduced([b * 1.0 if b is not None else None for b in elim._last[0]],
                                                  torch.nonzero(torch.cat([p.constraints.ravel() for p in asm.parts])
                                                                [elim.retained]).ravel()))
        longest = int((K.indptr[1:] - K.indptr[:-1]).max())
        print(json.dumps({"case": "tie+point", "block_size": elim.d, "n_dofs": asm.n_dofs, "n_retained": elim.n_retained,
                          "nnz_reduced": K.nnz, "longest_row": longest, "solve_ms_first": t1, "solve_ms_warm": t2,
                          "rigid_relation_max_err": float((u[1][top_b] - rigid).abs().max()),
                          "point_force": [round(v, 9) for v in f[2][0].tolist()],
                          "reduce_numeric_ms": t_numeric}), flush=True)
    except KeyError:
        pass
    except Exception:
        traceback.print_exc()


if __name__ == "__main__":
    main()
