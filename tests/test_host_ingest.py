"""Mesh ingest on the product path, no GPU: ADF-format CGNS reader with
the semantics of CgnsUnstructuredGrid::load, uniform refinement, INFO case files."""
import os

import numpy as np
import pytest

import oracle as O
from tests.adf_util import read_cgns_py, write_cgns
from tests.util import ORACLE_MESH_INT

REF_CGNS = "/root/reference/Examples/UnstructuredFlowAroundCylinder/case/CylinderMesh.cgns"


def host():
    from phase_b200.api import Communicator
    return Communicator(Communicator.HOST_ONLY)


def test_cgns_reader_on_a_written_file(tmp_path):
    from phase_b200.api import FiniteVolumeGrid2D as G
    # 2 quads + 2 triangles; BAR_2 boundary elements come FIRST (ids 1..4), like the shipped meshes
    xy = [[0, 0], [1, 0], [2, 0], [0, 1], [1, 1], [2, 1], [1, 2.0]]
    bars = [(0, 1), (1, 2), (2, 5), (3, 0)]
    quads = [(0, 1, 4, 3), (1, 2, 5, 4)]
    tris = [(3, 4, 6), (4, 5, 6)]
    path = str(tmp_path / "m.cgns")
    write_cgns(path, xy, [("BAR_2 1 - 4", 3, 1, bars), ("QUAD_4 5 - 6", 7, 5, quads), ("TRI_3 7 - 8", 5, 7, tris)],
               [("bottom", [1, 2]), ("sides", [3, 4])])
    g = G.from_cgns(host(), path)
    s = g.sizes()
    assert (s["nNodes"], s["nCells"], s["nPatches"]) == (7, 4, 2) and g.patch_names() == ["bottom", "sides"]
    # cells in element-id order, 0-based connectivity
    assert list(g.i32("cind")) == [0, 1, 4, 3, 1, 2, 5, 4, 3, 4, 6, 4, 5, 6]
    om = O.Mesh.create(xy, [0, 4, 8, 11, 14], g.i32("cind"))
    om.add_patch_by_nodes("bottom", [0, 1, 1, 2]); om.add_patch_by_nodes("sides", [2, 5, 3, 0])
    for k in ORACLE_MESH_INT:
        assert np.array_equal(g.i32(k), om.array(k)), k
    xy2, cptr2, cind2, patches2 = read_cgns_py(path)
    assert np.array_equal(cind2, g.i32("cind")) and [p[0] for p in patches2] == ["bottom", "sides"]


def test_cgns_reader_rejects_other_files(tmp_path):
    from phase_b200.api import FiniteVolumeGrid2D as G, PhaseB200Error
    p = tmp_path / "x.cgns"
    p.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\0" * 1000)
    with pytest.raises(PhaseB200Error, match="not an ADF-format"):
        G.from_cgns(host(), str(p))
    with pytest.raises(PhaseB200Error, match="cannot open"):
        G.from_cgns(host(), str(tmp_path / "missing.cgns"))


@pytest.mark.skipif(not os.path.exists(REF_CGNS), reason="reference example mesh not mounted")
def test_shipped_cylinder_mesh():
    from phase_b200.api import FiniteVolumeGrid2D as G
    g = G.from_cgns(host(), REF_CGNS)
    s = g.sizes()
    assert (s["nNodes"], s["nCells"]) == (7781, 15316)
    assert g.patch_names() == ["Cylinder", "TopBottom", "Inlet", "Outlet"]
    fp = g.i32("facePatch")
    assert list(np.bincount(fp[fp >= 0])) == [100, 82, 32, 32] and s["nBoundaryFaces"] == 246
    xy, cptr, cind, patches = read_cgns_py(REF_CGNS)
    om = O.Mesh.create(xy, cptr, cind)
    for name, pairs in patches:
        om.add_patch_by_nodes(name, pairs)
    for k in ORACLE_MESH_INT:
        assert np.array_equal(g.i32(k), om.array(k)), k
    assert np.allclose(g.f64("vol"), om.array("vol"), rtol=1e-12)


def test_uniform_refinement():
    from phase_b200.api import FiniteVolumeGrid2D as G
    g = G.triangulated(host(), 3, 2, 1.5, 1.0)
    r = g.refined(2)
    s0, s2 = g.sizes(), r.sizes()
    assert s2["nCells"] == 16 * s0["nCells"] and s2["nBoundaryFaces"] == 4 * s0["nBoundaryFaces"]
    assert np.isclose(r.f64("vol").sum(), 1.5) and r.f64("vol").min() > 0
    assert r.patch_names() == g.patch_names()
    fp0, fp2 = g.i32("facePatch"), r.i32("facePatch")
    assert np.array_equal(4 * np.bincount(fp0[fp0 >= 0]), np.bincount(fp2[fp2 >= 0]))
    # every boundary face is patched, none interior
    assert ((fp2 >= 0) == (r.i32("faceR") < 0)).all()
    q = G.rectilinear(host(), 2, 2).refined(1)
    assert q.sizes()["nCells"] == 16 and np.allclose(q.f64("vol"), 1.0 / 16)


@pytest.mark.skipif(not os.path.exists(REF_CGNS), reason="/root/reference absent")
def test_cylinder_fixture_is_the_shipped_mesh():
    """tests/golden/ref_cylinder_mesh.npz (what tools/cylinder_case.py runs config 3 from on the GPU box) rebuilds the mesh
    the reader makes of the shipped CGNS file: same faces, patches and link order."""
    from phase_b200.api import FiniteVolumeGrid2D as G
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cylinder_mesh.npz"))
    c = host()
    ref = G.from_cgns(c, REF_CGNS)
    g = G.from_cells(c, d["xy"], np.arange(0, 3 * len(d["tris"]) + 1, 3), d["tris"].ravel())
    for name in d["patch_order"]:
        g.createPatchByNodes(str(name), d["patch_" + str(name)].ravel())
    g.finalize()
    assert g.patch_names() == ref.patch_names() == ["Cylinder", "TopBottom", "Inlet", "Outlet"]
    for k in ("faceL", "faceR", "facePatch", "faceN1", "faceN2", "ilPtr", "ilCell", "ilFace", "blPtr", "blFace"):
        assert np.array_equal(ref.i32(k), g.i32(k)), k
    for k in ("vol", "faceSx", "faceSy"):
        assert np.array_equal(ref.f64(k), g.f64(k)), k
