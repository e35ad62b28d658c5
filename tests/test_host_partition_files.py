"""PhasePartitionGrid's per-partition grid files (U/utilities/PhasePartitionGrid.cpp:56-153) and the METIS option
(host integer work, no GPU): the C++ layout against a Python restatement of the reference's loops, and the written
ADF-CGNS file read back by the library's own CGNS reader and by an independent parser."""
import os

import numpy as np
import pytest

from phase_b200.adf import write_partition_grid
from phase_b200.api import Communicator, FiniteVolumeGrid2D as G, PhaseB200Error


@pytest.fixture(scope="module")
def host():
    c = Communicator(Communicator.HOST_ONLY)
    yield c
    c.close()


def reference_layout(g, part, proc, width):
    """PhasePartitionGrid.cpp:56-127 restated over the mesh tables"""
    ilPtr, ilCell, dlPtr, dlCell = g.i32("ilPtr"), g.i32("ilCell"), g.i32("dlPtr"), g.i32("dlCell")
    cptr, cind = g.i32("cptr"), g.i32("cind")
    cx, cy = g.f64("cellCx"), g.f64("cellCy")
    N = len(part)
    links = lambda c: list(ilCell[ilPtr[c]:ilPtr[c + 1]]) + list(dlCell[dlPtr[c]:dlPtr[c + 1]])
    cells, owner, bnd = [], [], []
    for c in range(N):
        if part[c] == proc:
            cells.append(c); owner.append(proc)
            if any(part[nb] != proc for nb in links(c)):
                bnd.append(c)
    seen = set()
    for c in bnd:
        for nb in links(c):
            if part[nb] != proc and nb not in seen:
                seen.add(nb); cells.append(nb); owner.append(part[nb])
        if width > 0:
            for k in range(N):
                if (cx[k] - cx[c]) ** 2 + (cy[k] - cy[c]) ** 2 <= width * width and part[k] != proc and k not in seen:
                    seen.add(k); cells.append(k); owner.append(part[k])
    g2l, nodes, eptr, eind = {}, [], [0], []
    for c in cells:
        for nd in cind[cptr[c]:cptr[c + 1]]:
            if nd not in g2l:
                g2l[nd] = len(nodes); nodes.append(nd)
            eind.append(g2l[nd] + 1)
        eptr.append(len(eind))
    return np.array(cells), np.array(owner), np.array(nodes), np.array(eptr), np.array(eind), g2l


@pytest.mark.parametrize("kind,nx,ny,nparts,width", [("rect", 10, 8, 3, 0.0), ("rect", 12, 12, 4, 0.2), ("tri", 6, 5, 2, 0.0)])
def test_partition_file_layout(host, kind, nx, ny, nparts, width):
    g = (G.rectilinear if kind == "rect" else G.triangulated)(host, nx, ny, 1.0, 1.0)
    part = g.partition_rcb(nparts)
    x, y = g.f64("nodeX"), g.f64("nodeY")
    n1, n2, fr, fp = g.i32("faceN1"), g.i32("faceN2"), g.i32("faceR"), g.i32("facePatch")
    names = g.patch_names()
    for proc in range(nparts):
        pf = g.partition_file(part, proc, width)
        cells, owner, nodes, eptr, eind, g2l = reference_layout(g, part, proc, width)
        assert np.array_equal(pf["GlobalID"], cells) and np.array_equal(pf["ProcNo"], owner)
        assert np.array_equal(pf["eptr"], eptr) and np.array_equal(pf["eind"], eind)
        assert np.array_equal(pf["nodes"], np.stack([x[nodes], y[nodes]], 1))
        for p, name in enumerate(names):
            want = []
            for f in np.flatnonzero((fr < 0) & (fp == p)):
                if n1[f] in g2l and n2[f] in g2l:
                    want += [g2l[n1[f]] + 1, g2l[n2[f]] + 1]
            assert list(pf["patches"].get(name, [])) == want, name
    g.close()


def test_partition_file_round_trip_through_cgns(host, tmp_path):
    from tests.adf_util import read_cgns_py
    g = G.rectilinear(host, 9, 7, 1.0, 1.0)
    part = g.partition_rcb(3)
    pf = g.partition_file(part, 1, 0.0)
    path = os.path.join(tmp_path, "Grid.cgns")
    write_partition_grid(path, "Cavity", pf)
    loc = G.from_cgns(host, path)                      # the library's own ADF-CGNS reader
    s = loc.sizes()
    assert s["nCells"] == len(pf["GlobalID"]) and s["nNodes"] == len(pf["nodes"])
    assert np.array_equal(loc.i32("cptr"), pf["eptr"]) and np.array_equal(loc.i32("cind"), pf["eind"] - 1)
    assert np.array_equal(np.stack([loc.f64("nodeX"), loc.f64("nodeY")], 1), pf["nodes"])
    assert sorted(loc.patch_names()) == sorted(pf["patches"])
    loc.close(); g.close()
    # GlobalID / ProcNo as an independent ADF parse sees them
    d = open(path, "rb").read()
    for name, want in (("GlobalID", pf["GlobalID"]), ("ProcNo", pf["ProcNo"])):
        off = d.find(b"NoDe" + name.encode().ljust(32))
        assert off > 0
        ptr = lambda b: int(b[:8], 16) * 4096 + int(b[8:12], 16)
        data = ptr(d[off + 230:off + 242])
        assert np.array_equal(np.frombuffer(d, np.int32, len(want), data + 16), want)


@pytest.mark.parametrize("method", ["mesh_dual", "graph_recursive"])
def test_metis_partition(host, method):
    g = G.rectilinear(host, 16, 16, 1.0, 1.0)
    try:
        part, cut = g.partition_metis(4, method)
    except PhaseB200Error as e:
        if e.code == -6:
            pytest.skip("built without METIS")
        raise
    counts = np.bincount(part, minlength=4)
    assert counts.min() >= 48 and counts.max() <= 80 and cut <= 40       # four compact quadrants cut 32 edges
    # the partition vector feeds the same local-mesh construction as any other
    loc = g.local(part, Communicator(Communicator.HOST_ONLY, 0, 4))
    assert loc.sizes()["nCells"] >= counts[0]
    loc.close(); g.close()
