"""Can the batched derivative evaluation of a hyperelastic energy be captured into a CUDA graph? (diagnostic)"""
import os
import sys
import time
import traceback

import torch
from torch.func import jacrev, vmap

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
torch.set_default_dtype(torch.float64)
torch.set_default_device("cuda")
from torchfem_b200.materials import small_matrix_mode  # noqa: E402


def psi(F, params):
    Cg = F.transpose(-1, -2) @ F
    logJ = 0.5 * torch.logdet(Cg)
    return params[0] / 2 * (torch.trace(Cg) - 3.0) - params[0] * logJ + params[1] / 2 * logJ ** 2


n = 131072
F = torch.eye(3).expand(n, 3, 3) + 0.05 * torch.randn(n, 3, 3)
params = torch.tensor([384.6, 576.9]).expand(n, 2).contiguous()


def run():
    with torch.enable_grad(), small_matrix_mode(F):
        P = vmap(jacrev(psi))(F, params)
        T = vmap(jacrev(jacrev(psi)))(F, params)
    return P.detach(), T.detach()


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


print("eager ms per evaluation:", timed(run))
try:
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        run()
        g.capture_begin()
        try:
            P, T = run()
        finally:
            g.capture_end()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    print("captured; replay ms:", timed(g.replay))
    Pe, Te = run()
    print("max diff P", float((P - Pe).abs().max()), "T", float((T - Te).abs().max()))
except Exception:
    traceback.print_exc()
