"""Host logic of `torchfem_b200.assembly` on the CPU: the elimination map T, the retained DOFs and the rigid modes
against fixtures from the unmodified reference (`oracle/make_golden.py::assembly_cases`, reference
assembly.py:180-358), and the reference's error behaviour (reference tests/test_assembly.py:274-322, 395-412, 458-461).
The parts are stand-ins that carry only what the coupling algebra reads (nodes and DOF counts): the models themselves
need a CUDA device (tests/test_gpu_assembly.py runs the solves)."""
import numpy as np
import pytest
import torch

from conftest import load_case


class _Part:
    """What `Assembly.coupling` / `_build_T` / `_rigid_modes` read of a model."""

    def __init__(self, nodes, n_dof_per_node):
        self.nodes = nodes
        self.n_nod = len(nodes)
        self.n_dof_per_node = n_dof_per_node
        self.n_dofs = self.n_nod * n_dof_per_node
        self._constraints = torch.zeros(self.n_nod, n_dof_per_node, dtype=torch.bool)


@pytest.fixture(scope="module")
def A():
    import torchfem_b200.assembly as A

    return A


@pytest.fixture(scope="module")
def gold():
    return load_case("assembly.npz")


def _check(asm, gold, tag):
    T, retained = asm._build_T()
    assert np.array_equal(T._indices().numpy(), gold[f"{tag}.T_idx"])
    assert np.abs(T._values().numpy() - gold[f"{tag}.T_val"]).max() <= 1e-15
    assert np.array_equal(retained.numpy(), gold[f"{tag}.retained"])
    assert np.abs(asm._rigid_modes().numpy() - gold[f"{tag}.modes"]).max() <= 1e-15


def test_tie_between_two_solids(A, gold):
    from torchfem_b200.mesh import cube_hexa

    n_a, _ = cube_hexa(4, 4, 3, 1.0, 1.0, 1.0)
    n_b, _ = cube_hexa(4, 4, 4, 1.0, 1.0, 1.0)
    n_b = n_b + torch.tensor([0.0, 0.0, 1.0])
    a, b = _Part(n_a, 3), _Part(n_b, 3)
    asm = A.Assembly([a, b])
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    _check(asm, gold, "tie")
    assert asm.n_dofs == a.n_dofs + b.n_dofs and asm.offsets == [0, a.n_dofs]


def test_reference_point_drives_a_face(A, gold):
    from torchfem_b200.mesh import cube_hexa

    nodes, _ = cube_hexa(4, 4, 4)
    solid, point = _Part(nodes, 3), A.ReferencePoint([0.5, 0.5, 2.0])
    assert point.n_dofs == 6 and point.forces.shape == (1, 6) and point.constraints.dtype == torch.bool
    asm = A.Assembly([solid, point])
    asm.coupling(solid, nodes[:, 2] == 1.0, point)
    _check(asm, gold, "point")
    # T reproduces u_s = u_p + theta x (x_s - x_p) for any point motion
    T, retained = asm._build_T()
    q = torch.zeros(len(retained))
    motion = torch.tensor([0.3, -0.2, 0.7, 0.05, -0.04, 0.02])
    q[-6:] = motion
    u = torch.sparse.mm(T, q[:, None])[:, 0][: solid.n_dofs].reshape(-1, 3)
    top = nodes[:, 2] == 1.0
    expected = motion[:3] + torch.cross(motion[3:].expand(int(top.sum()), 3), nodes[top] - point.nodes[0], dim=-1)
    assert torch.allclose(u[top], expected, atol=1e-15)


def test_subset_of_dofs(A, gold):
    from torchfem_b200.mesh import cube_hexa

    nodes, _ = cube_hexa(4, 4, 4)
    solid, point = _Part(nodes, 3), A.ReferencePoint([0.5, 0.5, 2.0])
    asm = A.Assembly([solid, point])
    asm.coupling(solid, nodes[:, 2] == 1.0, point, dofs=[2])
    _check(asm, gold, "subset")


def test_quadratic_part_tied_to_a_linear_one(A, gold):
    from torchfem_b200.elements import linear_to_quadratic
    from torchfem_b200.mesh import cube_hexa

    n_q, _ = linear_to_quadratic(*cube_hexa(3, 3, 3))
    n_l, _ = cube_hexa(3, 3, 3)
    n_l = n_l + torch.tensor([0.0, 0.0, 1.0])
    q, l = _Part(n_q, 3), _Part(n_l, 3)
    asm = A.Assembly([q, l])
    asm.coupling(l, n_l[:, 2] == 1.0, q, n_q[:, 2] == 1.0)
    _check(asm, gold, "mixed")


def test_planar_tie_and_point_with_one_rotation(A, gold):
    from torchfem_b200.mesh import rect_quad

    n_a, _ = rect_quad(4, 4, 1.0, 1.0)
    n_b, _ = rect_quad(4, 4, 1.0, 1.0)
    n_b = n_b + torch.tensor([1.0, 0.0])
    pa, pb, pp = _Part(n_a, 2), _Part(n_b, 2), A.ReferencePoint([2.5, 0.5])
    assert pp.n_dofs == 3
    asm = A.Assembly([pa, pb, pp])
    asm.coupling(pb, n_b[:, 0] == 1.0, pa, n_a[:, 0] == 1.0)
    asm.coupling(pb, n_b[:, 0] == 2.0, pp)
    _check(asm, gold, "planar")


def test_thermal_point(A):
    hp = A.ReferencePointHeat([0.5, 0.5, 2.5])
    assert hp.n_dofs == 1 and hp.heat_flux.shape == (1, 1) and hp.temperatures.shape == (1, 1)
    assert "thermal reference point" in repr(hp)
    asm = A.Assembly([hp])
    assert asm._rigid_modes().shape == (1, 1)


def _two_blocks():
    from torchfem_b200.mesh import cube_hexa

    n_a, _ = cube_hexa(2, 2, 2)
    n_b = n_a + torch.tensor([0.0, 0.0, 1.0])
    return _Part(n_a, 3), _Part(n_b, 3), n_a, n_b


def test_errors_follow_the_reference(A):
    a, b, n_a, n_b = _two_blocks()
    with pytest.raises(ValueError, match="only once"):
        A.Assembly([a, a])
    with pytest.raises(ValueError, match="one spatial dimension"):
        A.Assembly([a, A.ReferencePoint([0.0, 0.0])])
    with pytest.raises(ValueError, match="mechanical or thermal"):
        A.Assembly([a, A.ReferencePointHeat([0.0, 0.0, 1.0])])
    with pytest.raises(ValueError, match="must belong to this assembly"):
        A.Assembly([a]).coupling(a, n_a[:, 2] == 1.0, b)
    asm = A.Assembly([a, b])
    with pytest.raises(ValueError, match=r"indices in \[0, 3\)"):
        asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0, dofs=[5])
    point = A.ReferencePoint([0.5, 0.5, 1.0])
    with pytest.raises(ValueError, match="primary part that has none"):
        A.Assembly([a, point]).coupling(point, torch.ones(1, dtype=torch.bool), a, n_a[:, 2] == 1.0, dofs=[3])
    # a DOF eliminated twice
    asm = A.Assembly([a, b])
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    with pytest.raises(ValueError, match="more than one constraint"):
        asm._build_T()
    # a primary that is itself eliminated
    asm = A.Assembly([a, b, point])
    asm.coupling(a, n_a[:, 2] == 1.0, point)
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    with pytest.raises(ValueError, match="both eliminated and used as a primary"):
        asm._build_T()


def test_nearest_node_pairing(A):
    """Each secondary node follows the primary node it sits on (reference tests/test_assembly.py:325-342)."""
    a, b, n_a, n_b = _two_blocks()
    asm = A.Assembly([a, b])
    perm = torch.tensor([3, 1, 0, 2])
    top = torch.nonzero(n_a[:, 2] == 1.0).ravel()
    asm.coupling(b, n_b[:, 2] == 1.0, a, n_a[:, 2] == 1.0)
    (_, sec), (_, pri) = asm._links[0]
    assert torch.equal(n_b[sec][:, :2], n_a[pri][:, :2])
    # the candidate order does not matter
    asm2 = A.Assembly([a, b])
    mask = torch.zeros(len(n_a), dtype=torch.bool)
    mask[top[perm]] = True
    asm2.coupling(b, n_b[:, 2] == 1.0, a, mask)
    assert torch.equal(asm2._links[0][1][1], pri)
