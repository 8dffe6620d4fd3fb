"""Multi-GPU parity as driver-run tests (`-m gpu`): every case spawns `torchrun` on the visible GPUs and runs
`tools/multi_gpu_check.py`, which asserts fused peer-memory CG == host-driven NCCL CG == one-GPU solve of the whole
problem (<= 1e-9 relative), equal iteration counts (+-1) and a bitwise identical re-run. Skipped below 2 GPUs.
Covers the three halo layouts of SURVEY §8(e): Hexa1 slabs with range halos, forced index-list halos, and the Hexa2
coordinate partition of BASELINE configs[2]; `tools/damg_check.py` does the same for the distributed AMG-PCG."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_GPUS = torch.cuda.device_count() if torch.cuda.is_available() else 0
needs2 = pytest.mark.skipif(N_GPUS < 2, reason="needs at least 2 GPUs")
_PORT = [29600]


def _torchrun(world, script, *args, timeout=600):
    _PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]), os.path.join(ROOT, "tools", script), *args]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0, f"{' '.join(cmd)}\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    assert lines, r.stdout[-2000:]
    return lines[-1]


def _worlds():
    return sorted({2, N_GPUS} - {0, 1}) if N_GPUS >= 2 else [2]


@needs2
@pytest.mark.parametrize("world", _worlds())
def test_fused_cg_hexa1_range_halos(world):
    out = _torchrun(world, "multi_gpu_check.py", "--edge", "16")
    assert out["ok"] and out["bitwise_reproducible"] and out["rel_fused_vs_single"] <= 1e-9
    assert out["iters_fused"] == out["iters_nccl"]


@needs2
def test_fused_cg_index_list_halos():
    out = _torchrun(2, "multi_gpu_check.py", "--edge", "12", "--general")
    assert out["ok"] and out["general_halo"] and out["rel_fused_vs_nccl"] <= 1e-9


@needs2
@pytest.mark.parametrize("world", _worlds())
def test_fused_cg_hexa2_coordinate_partition(world):
    out = _torchrun(world, "multi_gpu_check.py", "--edge", "6", "--hexa2")
    assert out["ok"] and out["hexa2"] and out["rel_fused_vs_single"] <= 1e-9


@needs2
@pytest.mark.parametrize("world", _worlds())
def test_distributed_amg_pcg(world):
    if not os.path.exists(os.path.join(ROOT, "tools", "damg_check.py")):
        pytest.skip("distributed AMG not built")
    out = _torchrun(world, "damg_check.py", "--edge", "16")
    assert out["ok"], out
