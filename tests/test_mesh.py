"""Mesh generators against the reference's (run here when /root/reference exists) and against the oracle's
restatement; shapes/counts as in reference tests/test_mesh.py:91-115."""
import numpy as np
import pytest
import torch

from oracle import fem_oracle as O
from oracle import ref_import


@pytest.fixture(scope="module")
def M():
    torch.set_default_device("cpu")
    from torchfem_b200 import mesh

    return mesh


def test_cube_hexa_matches_oracle_and_counts(M):
    n, e = M.cube_hexa(5, 4, 3, 2.0, 1.0, 0.5)
    n_ref, e_ref = O.cube_hexa(5, 4, 3, 2.0, 1.0, 0.5)
    assert np.array_equal(e.numpy(), e_ref) and np.allclose(n.numpy(), n_ref, atol=0)
    assert n.shape == (60, 3) and e.shape == (24, 8) and e.dtype == torch.int64
    assert float(n[:, 0].max()) == 2.0 and float(n.min()) == 0.0


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present (GPU box)")
def test_all_generators_identical_to_reference(M):
    tf = ref_import.load()
    from torchfem import mesh as R

    cases = [("cube_hexa", (4, 3, 5, 1.0, 2.0, 0.5)), ("cube_tetra", (4, 3, 3)), ("rect_quad", (5, 4, 2.0, 1.0)),
             ("rect_tri", (4, 5)), ("rect_tri", (4, 5, 1.0, 1.0, "up")), ("rect_tri", (4, 5, 1.0, 1.0, "down")),
             ("rect_tri", (3, 3, 1.0, 1.0, "center"))]
    for name, args in cases:
        n1, e1 = getattr(M, name)(*args)
        n2, e2 = getattr(R, name)(*args)
        assert torch.equal(e1, e2), (name, args)
        assert torch.equal(n1, n2), (name, args)
    with pytest.raises(ValueError, match="Unknown variant"):
        M.rect_tri(3, 3, variant="nope")
