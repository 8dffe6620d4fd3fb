"""Where does `Solid.solve` spend its time? torch.profiler over one solve of the cube (config B by default).
    python tools/prof_api.py --edge 150
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--edge", type=int, default=150)
    ap.add_argument("--rows", type=int, default=45)
    ap.add_argument("--backward", action="store_true", help="forces as differentiable parameter + u.sum().backward()")
    ap.add_argument("--stol", type=float, default=1e-8)
    ap.add_argument("--method", default="cg")
    a = ap.parse_args()
    torch.set_default_dtype(torch.float64)
    import torchfem_b200 as T
    from torchfem_b200.materials import IsotropicElasticity3D
    from torchfem_b200.mesh import cube_hexa

    dev = torch.device("cuda", 0)
    with torch.device("cpu"):
        nodes, elements = cube_hexa(a.edge + 1, a.edge + 1, a.edge + 1)
    n_elem = len(elements)
    model = T.Solid(nodes.to(dev), elements.to(dev),
                    IsotropicElasticity3D(torch.full((n_elem,), 1000.0, device=dev), torch.full((n_elem,), 0.3, device=dev)))
    con = torch.zeros_like(model.nodes, dtype=torch.bool)
    disp = torch.zeros_like(model.nodes)
    con[model.nodes[:, 0] == 0.0, :] = True
    con[model.nodes[:, 0] == 1.0, 0] = True
    disp[model.nodes[:, 0] == 1.0, 0] = 0.1
    model.constraints, model.displacements = con, disp

    if a.backward:
        model.forces = torch.zeros_like(model.nodes, requires_grad=True)

    def run():
        torch.cuda.synchronize()
        t = time.perf_counter()
        if a.backward:
            u, *_ = model.solve(method=a.method, stol=a.stol, differentiable_parameters=model.forces)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            u.sum().backward()
            torch.cuda.synchronize()
            print("   fwd %.3f s  bwd %.3f s" % (t1 - t, time.perf_counter() - t1))
        else:
            u, *_ = model.solve(method=a.method, stol=a.stol, rtol=1e-6)
        torch.cuda.synchronize()
        return time.perf_counter() - t

    print("warm-up solve: %.3f s" % run())
    print("second solve:  %.3f s" % run())
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        run()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=a.rows, max_name_column_width=70))


if __name__ == "__main__":
    main()
