"""ADF (classic CGNS container) writer and the per-partition grid files of the reference's PhasePartitionGrid utility.

Layout as parsed by phase_b200/csrc/ingest.cu and by cgnslib: file offset = block*4096 + offset;
node = "NoDe" name[32] label[32] nSub(8 hex) nEntries(8) subPtr(8+4 hex) dtype[32] nDims(2 hex) 12 x dim(8 hex)
nChunks(4 hex) dataPtr(8+4 hex) "TaiL" (246 bytes); sub-node table = "SNTb" endPtr(12) + entries name[32] ptr(12);
data = "DaTa" endPtr(12) payload.
"""
import numpy as np

BAR_2, TRI_3, QUAD_4, MIXED = 3, 5, 7, 20


def _ptr(off):
    return ("%08x%04x" % (off // 4096, off % 4096)).encode()


class AdfWriter:
    def __init__(self):
        hdr = bytearray(b"\xc0\xa8\xa3\xa9ADF Database Version A02011>AdF0")
        hdr += b" " * (266 - len(hdr))
        self.buf = hdr
        self.nodes = []

    def _alloc(self, n):
        off = len(self.buf)
        self.buf += b"\0" * n
        return off

    def add(self, name, label, dtype="MT", data=None, dims=None, children=()):
        """children: list of node offsets already written. returns this node's offset."""
        data_off = 4096
        nch = 0
        if data is not None:
            payload = data if isinstance(data, bytes) else np.ascontiguousarray(data).tobytes()
            data_off = self._alloc(16 + len(payload) + 4)
            self.buf[data_off:data_off + 4] = b"DaTa"
            self.buf[data_off + 4:data_off + 16] = _ptr(data_off + 16 + len(payload))
            self.buf[data_off + 16:data_off + 16 + len(payload)] = payload
            self.buf[data_off + 16 + len(payload):data_off + 20 + len(payload)] = b"dEnD"
            nch = 1
        sub_off = 0
        if children:
            sub_off = self._alloc(16 + 44 * len(children) + 4)
            self.buf[sub_off:sub_off + 4] = b"SNTb"
            self.buf[sub_off + 4:sub_off + 16] = _ptr(sub_off + 16 + 44 * len(children))
            p = sub_off + 16
            for cname, coff in children:
                self.buf[p:p + 32] = cname.encode().ljust(32)
                self.buf[p + 32:p + 44] = _ptr(coff)
                p += 44
        dims = list(dims or [])
        off = self._alloc(246)
        rec = bytearray(b"NoDe")
        rec += name.encode().ljust(32) + label.encode().ljust(32)
        rec += ("%08x" % len(children)).encode() + ("%08x" % len(children)).encode() + _ptr(sub_off)
        rec += dtype.encode().ljust(32) + ("%02x" % len(dims)).encode()
        for i in range(12):
            rec += ("%08x" % (dims[i] if i < len(dims) else 0)).encode()
        rec += ("%04x" % nch).encode() + _ptr(data_off) + b"TaiL"
        assert len(rec) == 246, len(rec)
        self.buf[off:off + 246] = rec
        return off

    def finish(self, path, root_children):
        # the reader takes the FIRST "NoDe" in the file as the root: write the root at offset 266
        root = bytearray(b"NoDe") + b"ADF MotherNode".ljust(32) + b"Root Node of ADF File".ljust(32)
        sub_off = self._alloc(16 + 44 * len(root_children) + 4)
        self.buf[sub_off:sub_off + 4] = b"SNTb"
        p = sub_off + 16
        for cname, coff in root_children:
            self.buf[p:p + 32] = cname.encode().ljust(32)
            self.buf[p + 32:p + 44] = _ptr(coff)
            p += 44
        root += ("%08x" % len(root_children)).encode() * 2 + _ptr(sub_off) + b"MT".ljust(32) + b"00" + b"00000000" * 12
        root += b"0000" + _ptr(4096) + b"TaiL"
        assert len(root) == 246
        self.buf[266:266 + 246] = root
        open(path, "wb").write(bytes(self.buf))



def write_partition_grid(path, case_name, pf):
    """solution/Proc<k>/Grid.cgns as U/utilities/PhasePartitionGrid.cpp:128-155 lays it out: base (2, 2) named after
    the case, one unstructured zone "Zone", coordinates, a MIXED section "Cells" over elements 1..nCells, one BAR_2
    section + BC (PointRange, EdgeCenter) per patch, and the cell-centred solution "Info" with the integer fields
    GlobalID and ProcNo.  `pf` = FiniteVolumeGrid2D.partition_file(...)."""
    w = AdfWriter()
    w._alloc(512)  # room for the root record at 266
    xy = np.asarray(pf["nodes"], float)
    n_cells = len(pf["GlobalID"])
    cx = w.add("CoordinateX", "DataArray_t", "R8", xy[:, 0].copy(), [len(xy)])
    cy = w.add("CoordinateY", "DataArray_t", "R8", xy[:, 1].copy(), [len(xy)])
    gc = w.add("GridCoordinates", "GridCoordinates_t", children=[("CoordinateX", cx), ("CoordinateY", cy)])
    zt = w.add("ZoneType", "ZoneType_t", "C1", b"Unstructured", [12])
    kids = [("ZoneType", zt), ("GridCoordinates", gc)]
    eptr, eind = pf["eptr"], pf["eind"]
    conn = []
    for i in range(n_cells):
        k = eptr[i + 1] - eptr[i]
        conn.append({2: BAR_2, 3: TRI_3, 4: QUAD_4}[int(k)])
        conn.extend(int(v) for v in eind[eptr[i]:eptr[i + 1]])
    conn = np.asarray(conn, np.int32)
    er = w.add("ElementRange", "IndexRange_t", "I4", np.array([1, n_cells], np.int32), [2])
    ec = w.add("ElementConnectivity", "DataArray_t", "I4", conn, [len(conn)])
    kids.append(("Cells", w.add("Cells", "Elements_t", "I4", np.array([MIXED, 0], np.int32), [2],
                                children=[("ElementRange", er), ("ElementConnectivity", ec)])))
    end, bck = n_cells, []
    for name, pairs in pf["patches"].items():
        start = end + 1
        end = start + len(pairs) // 2 - 1
        er = w.add("ElementRange", "IndexRange_t", "I4", np.array([start, end], np.int32), [2])
        ec = w.add("ElementConnectivity", "DataArray_t", "I4", np.asarray(pairs, np.int32), [len(pairs)])
        kids.append((name, w.add(name, "Elements_t", "I4", np.array([BAR_2, 0], np.int32), [2],
                                 children=[("ElementRange", er), ("ElementConnectivity", ec)])))
        pr = w.add("PointRange", "IndexRange_t", "I4", np.array([start, end], np.int32), [1, 2])
        gl = w.add("GridLocation", "GridLocation_t", "C1", b"EdgeCenter", [10])
        bck.append((name, w.add(name, "BC_t", "C1", b"BCGeneral", [9], children=[("PointRange", pr), ("GridLocation", gl)])))
    kids.append(("ZoneBC", w.add("ZoneBC", "ZoneBC_t", children=bck)))
    gl = w.add("GridLocation", "GridLocation_t", "C1", b"CellCenter", [10])
    gid = w.add("GlobalID", "DataArray_t", "I4", np.asarray(pf["GlobalID"], np.int32), [n_cells])
    pno = w.add("ProcNo", "DataArray_t", "I4", np.asarray(pf["ProcNo"], np.int32), [n_cells])
    kids.append(("Info", w.add("Info", "FlowSolution_t", children=[("GridLocation", gl), ("GlobalID", gid), ("ProcNo", pno)])))
    zone = w.add("Zone", "Zone_t", "I4", np.array([len(xy), n_cells, 0], np.int32), [1, 3], children=kids)
    base = w.add(case_name, "CGNSBase_t", "I4", np.array([2, 2], np.int32), [2], children=[("Zone", zone)])
    ver = w.add("CGNSLibraryVersion", "CGNSLibraryVersion_t", "R4", np.array([3.1], np.float32), [1])
    w.finish(path, [("CGNSLibraryVersion", ver), (case_name, base)])
