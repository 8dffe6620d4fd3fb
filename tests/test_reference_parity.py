"""The oracle against the UNMODIFIED reference run here (`/root/reference/src` in the build container, the staged
`oracle/_ref` on the GPU box; skipped when neither exists): the timed CPU flow of `bench.py --impl reference`
(oracle/ref_bench.py) must give the oracle port's vectors, and the staged copy must be the reference's own files."""
import hashlib
import os

import numpy as np
import pytest

from oracle import fem_oracle as O
from oracle import ref_bench as R
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference package not available")


def test_reference_flow_equals_oracle_port_and_golden():
    E = 10
    model, _ = R.cube_extension_model(E)
    u, ph = R.linear_solve(model, rtol=1e-10)
    nodes, elements = O.cube_hexa(E + 1, E + 1, E + 1)
    bref, w = O.hexa1_tables()
    con_mask, disp = O.cube_extension_bcs(nodes)
    out = O.linear_solve_reference_flow(nodes, elements, bref, w, O.isotropic_C3d(1000.0, 0.3, len(elements)), con_mask,
                                        disp, rtol=1e-10)
    assert abs(ph["iterations"] - out["iterations"]) <= 1
    assert np.linalg.norm(u - out["u"]) <= 1e-9 * np.linalg.norm(out["u"])
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "config_a.npz")))
    assert np.linalg.norm(u - g["u"]) <= 1e-8 * np.linalg.norm(g["u"])
    assert ph["n_dofs"] == 3993 and R.hot_path_seconds(ph) > 0.0


def test_staged_copy_is_the_unmodified_reference():
    staged = os.path.join(os.path.dirname(os.path.abspath(O.__file__)), "_ref", "torchfem")
    source = "/root/reference/src/torchfem"
    if not (os.path.isdir(staged) and os.path.isdir(source)):
        pytest.skip("needs both the staged copy and /root/reference")
    for name in ("base.py", "sparse.py", "elements.py", "solid.py", "mesh.py"):
        a = hashlib.sha256(open(os.path.join(staged, name), "rb").read()).hexdigest()
        b = hashlib.sha256(open(os.path.join(source, name), "rb").read()).hexdigest()
        assert a == b, name
