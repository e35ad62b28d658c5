"""PhasePartitionGrid (U/utilities/PhasePartitionGrid.cpp): partition a grid and write one grid file per partition,
solution/Proc<k>/Grid.cgns, with the GlobalID / ProcNo arrays the reference's solvers read back.

    python tools/partition_grid.py -n 4 [-m 0.0] [--mesh file.cgns | --rect NX NY W H] [--method graph_recursive|mesh_dual|rcb]
                                   [--out solution] [--case-name Case]

Host-only (no GPU needed): mesh build, METIS (the toolkit's libmetis_static.a) or RCB, the file layout
(phb_partition_file_build) and the ADF writer (phase_b200/adf.py).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("-n", "--num-partitions", type=int, required=True)
    ap.add_argument("-m", "--min-buffer-width", type=float, default=0.0)
    ap.add_argument("--mesh")
    ap.add_argument("--rect", nargs=4, metavar=("NX", "NY", "W", "H"))
    ap.add_argument("--method", default="graph_recursive", choices=["graph_recursive", "mesh_dual", "rcb"])
    ap.add_argument("--out", default="solution")
    ap.add_argument("--case-name", default="Case")
    a = ap.parse_args()
    from phase_b200.adf import write_partition_grid
    from phase_b200.api import Communicator, FiniteVolumeGrid2D as G
    host = Communicator(Communicator.HOST_ONLY)
    if a.mesh:
        g = G.from_cgns(host, a.mesh)
    else:
        nx, ny, w, h = a.rect or (16, 16, 1.0, 1.0)
        g = G.rectilinear(host, int(nx), int(ny), float(w), float(h))
    if a.method == "rcb":
        part, cut = g.partition_rcb(a.num_partitions), None
    else:
        part, cut = g.partition_metis(a.num_partitions, a.method)
    print("partitioned %d cells into %d parts (%s%s)" % (len(part), a.num_partitions, a.method,
                                                         "" if cut is None else ", edge cut %d" % cut))
    for proc in range(a.num_partitions):
        pf = g.partition_file(part, proc, a.min_buffer_width)
        d = os.path.join(a.out, "Proc%d" % proc)
        os.makedirs(d, exist_ok=True)
        write_partition_grid(os.path.join(d, "Grid.cgns"), a.case_name, pf)
        print("proc %d: %d cells (%d owned), %d nodes, patches %s" % (proc, len(pf["GlobalID"]), int((pf["ProcNo"] == proc).sum()),
                                                                      len(pf["nodes"]), sorted(pf["patches"])))
    g.close(); host.close()


if __name__ == "__main__":
    main()
