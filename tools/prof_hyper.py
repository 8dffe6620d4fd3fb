"""torch.profiler over the forward solve of the reference's hyperelasticity benchmark (N = 65, 10 increments)."""
import math
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
torch.set_default_device("cuda")
import torchfem_b200 as T  # noqa: E402
from torchfem_b200.materials import Hyperelastic3D  # noqa: E402
from torchfem_b200.mesh import cube_hexa  # noqa: E402

N, STRETCH = 65, 10.0
En, NU = 1000.0, 0.3
LBD = En * NU / ((1.0 + NU) * (1.0 - 2.0 * NU))
MU = En / (2.0 * (1.0 + NU))


def psi(F, params):
    Cg = F.transpose(-1, -2) @ F
    logJ = 0.5 * torch.logdet(Cg)
    return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2


lx = 4.0 / (N - 1)
nodes, elements = cube_hexa(5, N, N, lx, 1.0, 1.0)
params = torch.tensor([MU, LBD], requires_grad=True)
right = nodes[:, 0] == lx


def build():
    box = T.Solid(nodes, elements, Hyperelastic3D(psi, params))
    box.constraints[nodes[:, 0] == 0.0, 0] = True
    box.constraints[right, 0] = True
    box.constraints[nodes[:, 1] == 0.5, 1] = True
    box.constraints[nodes[:, 2] == 0.5, 2] = True
    box.displacements[right, 0] = (STRETCH - 1.0) * lx
    return box


lam = torch.logspace(0, math.log10(STRETCH), 11)
increments = (lam - 1.0) / (STRETCH - 1.0)


def run():
    box = build()
    torch.cuda.synchronize()
    t = time.perf_counter()
    box.solve(increments=increments, nlgeom=True, differentiable_parameters=params, method="cg", stol=1e-10)
    torch.cuda.synchronize()
    return time.perf_counter() - t


print("warm-up %.3f s" % run())
print("second  %.3f s" % run())
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    run()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))
