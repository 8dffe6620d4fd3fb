"""Transient heat conduction with method="amgx" vs "cg" (system matrix and hierarchy reused across the time steps)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
torch.set_default_device("cuda")
import torchfem_b200 as T  # noqa: E402
from torchfem_b200.materials import IsotropicConductivity3D  # noqa: E402
from torchfem_b200.mesh import cube_hexa  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 61
nodes, elements = cube_hexa(N, N, N)
out = {}
for method in ("cg", "amgx"):
    cube = T.SolidHeat(nodes, elements, IsotropicConductivity3D(kappa=2.0, rho=30.0))
    cube.constraints[nodes[:, 0] == 0.0] = True
    cube.temperatures[nodes[:, 0] == 0.0, 0] = 1.0
    cube.heat_flux[nodes[:, 0] == 1.0, 0] = 0.05
    for rep in range(2):
        torch.cuda.synchronize()
        t = time.perf_counter()
        temp, *_ = cube.time_integration(torch.tensor([0.5, 1.0, 2.0]), delta_t=0.25, method=method, stol=1e-11)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
    out[method] = temp
    print(method, "n_dofs", cube.n_dofs, "steps 8, %.3f s" % dt)
print("max rel diff", float((out["amgx"] - out["cg"]).abs().max() / out["cg"].abs().max()))
