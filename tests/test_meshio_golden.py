"""Row (f)1 of SURVEY.md §8: .msh ingestion and the golden-dump loader (CPU parts)."""
import numpy as np
import pytest

from common import OracleStepper, make_case, make_oracle
from shakti_b200 import golden, meshgen, meshio_lite


@pytest.mark.parametrize("version", [2, 4])
def test_msh_round_trip(tmp_path, version):
    xy, cells = meshgen.rectangle(7, 5, 7e3, 5e3, jitter=0.2, diagonal="random")
    f = tmp_path / "m.msh"
    meshio_lite.write_msh(f, xy, cells, version=version)
    domain, ctags, ftags = meshio_lite.read_from_msh(str(f), None, gdim=2)
    assert np.allclose(domain.geometry.x[:, :2], xy, rtol=0, atol=1e-12 * 7e3)
    assert np.array_equal(domain.cells, cells) and ftags is None and (ctags == 1).all()


def test_msh_keeps_only_physical_cells_and_used_nodes(tmp_path):
    f = tmp_path / "m.msh"
    f.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n$Nodes\n5\n1 0 0 0\n2 1 0 0\n3 0 1 0\n4 1 1 0\n9 5 5 0\n$EndNodes\n"
                 "$Elements\n4\n1 15 2 0 1 9\n2 1 2 0 1 1 2\n3 2 2 7 1 1 2 3\n4 2 2 0 2 2 4 3\n$EndElements\n")
    xy, cells, ctag = meshio_lite.read_msh_arrays(str(f))
    assert cells.tolist() == [[0, 1, 2]] and ctag.tolist() == [7] and xy.shape == (3, 2)


def test_msh_rejects_garbage(tmp_path):
    f = tmp_path / "x.msh"
    f.write_text("hello\n")
    with pytest.raises(ValueError):
        meshio_lite.read_msh_arrays(str(f))


def test_golden_dump_round_trip_with_oracle(tmp_path):
    c = make_case(nx=10, ny=8, seed=12)
    golden.write_dump(tmp_path / "dump", make_oracle(*c), [360.0, 3600.0, 3600.0])
    d = golden.Dump(tmp_path / "dump")
    assert len(d.steps) == 3 and d.cells.dtype == np.int32 and d.quad[1].sum() == pytest.approx(0.5)
    rep = golden.check(d, OracleStepper(d))
    assert rep["F"] == 0.0 and rep["J"] == 0.0 and rep["step2"]["N"] == 0.0
    # a different quadrature table must be detected through the K integral
    from oracle import quadrature
    d.quad = quadrature.gauss_jacobi_triangle(3)
    with pytest.raises(AssertionError):
        golden.check(d, OracleStepper(d))
