"""Writes tests/golden/ref_cylinder_mesh.npz: nodes, triangles and boundary patches of the reference's shipped
Examples/UnstructuredFlowAroundCylinder/case/CylinderMesh.cgns (config 3's mesh: 7781 nodes, 15 316 triangles, patches
Cylinder / TopBottom / Inlet / Outlet), read here with the ADF-CGNS reader (phb_mesh_read_cgns, bit-checked against the
oracle in tests/test_host_ingest.py).  The GPU box has no /root/reference; tools/cylinder_case.py rebuilds the mesh from
this fixture and refines it there (5 rounds = 15.7M cells).      python tests/golden/make_cylinder_mesh.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
REF = "/root/reference/Examples/UnstructuredFlowAroundCylinder/case/CylinderMesh.cgns"


def main():
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G
    c = Communicator(Communicator.HOST_ONLY)
    g = G.from_cgns(c, REF)
    xy = np.stack([g.f64("nodeX"), g.f64("nodeY")], 1)
    cptr, cind = g.i32("cptr"), g.i32("cind")
    assert np.all(np.diff(cptr) == 3)
    fr, fp, n1, n2 = g.i32("faceR"), g.i32("facePatch"), g.i32("faceN1"), g.i32("faceN2")
    out = {"xy": xy, "tris": cind.reshape(-1, 3).astype(np.int32)}
    names = g.patch_names()
    for k, name in enumerate(names):
        sel = np.nonzero((fr < 0) & (fp == k))[0]
        out["patch_" + name] = np.stack([n1[sel], n2[sel]], 1).astype(np.int32)
    out["patch_order"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ref_cylinder_mesh.npz"), **out)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
    g.close(); c.close()


if __name__ == "__main__":
    main()
