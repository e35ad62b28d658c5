"""I1-I4 on the product path (C ABI, host mesh build of libphase_b200) against the
oracle: integer artefacts bit-exact, geometry to round-off."""
import numpy as np
import pytest

import oracle as O
from tests.util import ORACLE_MESH_INT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def comm():
    from phase_b200.api import Communicator
    c = Communicator(0)
    yield c
    c.close()


@pytest.mark.parametrize("kind,nx,ny", [("rect", 3, 3), ("rect", 17, 9), ("tri", 6, 5), ("tri", 1, 1), ("rect", 1, 4)])
def test_connectivity_bit_exact(comm, kind, nx, ny):
    from phase_b200.api import FiniteVolumeGrid2D as G
    g = (G.rectilinear if kind == "rect" else G.triangulated)(comm, nx, ny, 1.3, 0.7)
    om = (O.Mesh.rectilinear if kind == "rect" else O.Mesh.triangulated)(nx, ny, 1.3, 0.7)
    for name in ORACLE_MESH_INT:
        assert np.array_equal(g.i32(name), om.array(name)), name
    # facePatch ids: both create x-, x+, y-, y+ in that order
    for a, b in (("vol", "vol"), ("cellCx", "cellCx"), ("cellCy", "cellCy"), ("faceCx", "faceCx"), ("faceCy", "faceCy")):
        assert np.allclose(g.f64(a), om.array(b), rtol=1e-14, atol=1e-15)
    # S_f oriented out of lCell == outward normal of the lCell link
    s = g.sizes()
    assert s["nFaces"] == om.sizes["nFaces"] and s["nnz"] == s["nCells"] + 2 * s["nInteriorFaces"]
    g.close()


def test_canonical_pattern_matches_reference_compact_layout(comm):
    from phase_b200.api import FiniteVolumeGrid2D as G
    g = G.triangulated(comm, 7, 4)
    om = O.Mesh.triangulated(7, 4)
    fs = O.cavity(om)
    rp, ci, va, rhs = fs.assemble_u(0.01).export()      # reference compact [P, nb...]
    N = om.sizes["nCells"]
    assert np.array_equal(g.i32("rowPtr"), rp[:N + 1])
    assert np.array_equal(g.i32("colInd"), ci[:rp[N]])
    # face -> slot map addresses exactly the (lCell,rCell) and (rCell,lCell) entries
    sl, sr, fl, fr = g.i32("slotL"), g.i32("slotR"), g.i32("faceL"), g.i32("faceR")
    col, rowp = g.i32("colInd"), g.i32("rowPtr")
    for f in range(len(fl)):
        if fr[f] < 0:
            assert sl[f] == -1 and sr[f] == -1
            continue
        assert col[sl[f]] == fr[f] and rowp[fl[f]] <= sl[f] < rowp[fl[f] + 1]
        assert col[sr[f]] == fl[f] and rowp[fr[f]] <= sr[f] < rowp[fr[f] + 1]
    g.close()


def test_generic_mesh_and_patches(comm):
    from phase_b200.api import FiniteVolumeGrid2D as G
    # two quads + one triangle, mixed; patch by node pairs
    xy = np.array([[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1], [1, 2.0]])
    cptr = [0, 4, 8, 11]
    cind = [0, 1, 4, 3, 1, 2, 5, 4, 3, 4, 6]
    g = G.from_cells(comm, xy, cptr, cind)
    om = O.Mesh.create(xy, cptr, cind)
    assert g.createPatchByNodes("bottom", [0, 1, 1, 2]) == om.add_patch_by_nodes("bottom", [0, 1, 1, 2])
    g.finalize()
    for name in ORACLE_MESH_INT:
        assert np.array_equal(g.i32(name), om.array(name)), name
    from phase_b200.api import PhaseB200Error
    g2 = G.from_cells(comm, xy, cptr, cind)
    with pytest.raises(PhaseB200Error):
        g2.createPatchByNodes("bad", [0, 6])     # findFace throws in the reference too
    g.close(); g2.close()
