"""Partition and halo maps (I3, I5) of the oracle against the REFERENCE'S OWN FiniteVolumeGrid2D::partition,
initCommBuffers and IndexMap (UG/FiniteVolumeGrid2D.cpp:276-392,459-511, UE/IndexMap.cpp:5-40), run inside
oracle/_ref/libphase_ref_fv.so with the MPI ranks as threads (oracle/ref_mpi_threads.cpp) and the partition vector
handed to the METIS hook: local cell sets and numbering, ownership, buffer and send groups (with their order), local
face connectivity and the global row numbers of owned and ghost cells -- all bit-exact."""
import numpy as np
import pytest

import oracle as O
from oracle import ref_fv as R

pytestmark = pytest.mark.skipif(not R.available(), reason="reference FV library not built (no /root/reference, no prebuilt .so)")


def partitions(nx, ny, P, kind):
    om = O.Mesh.rectilinear(nx, ny, 1.0, 1.0)
    cx, cy = om.array("cellCx"), om.array("cellCy")
    if kind == "strips":
        part = np.minimum((cy * P).astype(np.int32), P - 1)
    elif kind == "columns":
        part = np.minimum((cx * P).astype(np.int32), P - 1)
    elif kind == "blocks":
        px = 2
        part = np.minimum((cx * px).astype(np.int32), px - 1) + px * np.minimum((cy * (P // px)).astype(np.int32), P // px - 1)
    else:   # ragged: a diagonal cut with an island
        part = ((cx + 0.7 * cy) * P / 1.7).astype(np.int32).clip(0, P - 1)
        part[(np.abs(cx - 0.3) < 0.12) & (np.abs(cy - 0.6) < 0.12)] = P - 1
    return om, part.astype(np.int32)


@pytest.mark.parametrize("nx,ny,P,kind", [(8, 6, 2, "strips"), (8, 6, 3, "columns"), (12, 10, 4, "blocks"), (14, 12, 3, "ragged"),
                                          (9, 9, 8, "strips")])
def test_partition_and_halo_maps_match_reference(nx, ny, P, kind):
    om, part = partitions(nx, ny, P, kind)
    case = R.Case(nx, ny, 1.0, 1.0)
    ref = R.partition(case, part, P, n_indices=1)
    ref2 = R.partition(case, part, P, n_indices=2)
    locs = om.partition(part, P)
    n_local = [int((locs[r].array("owner") == r).sum()) for r in range(P)]
    offset = np.concatenate([[0], np.cumsum(n_local)])
    for r in range(P):
        for k in ("globalId", "owner", "bufPtr", "bufCell", "sendPtr", "sendCell", "faceL", "faceR"):
            assert np.array_equal(ref[r][k], locs[r].array(k)), (r, k)
        lr, gr, owner = locs[r].array("localRow"), locs[r].array("globalRow"), locs[r].array("owner")
        assert np.array_equal(ref[r]["local"], lr) and np.array_equal(ref[r]["global"], gr)
        # two index sets (vector equations): [x-block | y-block] per rank, ghosts carry their owner's numbers
        n = len(lr)
        in_owner = gr - offset[owner]
        for s in (0, 1):
            want_local = np.where(lr >= 0, lr + s * n_local[r], -1)
            want_global = 2 * offset[owner] + s * np.asarray(n_local)[owner] + in_owner
            assert np.array_equal(ref2[r]["local"][s * n:(s + 1) * n], want_local)
            assert np.array_equal(ref2[r]["global"][s * n:(s + 1) * n], want_global)
    case.close()
