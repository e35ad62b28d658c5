"""Minimal ADF (classic CGNS container) WRITER + independent reader, for tests only.
Layout as parsed by phase_b200/csrc/ingest.cu: file offset = block*4096 + offset;
node = "NoDe" name[32] label[32] nSub(8 hex) nEntries(8) subPtr(8+4 hex) dtype[32]
nDims(2 hex) 12 x dim(8 hex) nChunks(4 hex) dataPtr(8+4 hex) "TaiL" (246 bytes);
sub-node table = "SNTb" endPtr(12) + entries name[32] ptr(12); data = "DaTa" endPtr(12) payload."""
import numpy as np


from phase_b200.adf import AdfWriter, _ptr  # noqa: F401  (the writer is product code: partition files)


def write_cgns(path, xy, elements, bcs):
    """elements: list of (section name, type code, first id, list of node tuples 0-based);
    bcs: list of (name, [element ids]).  Element type codes: BAR_2 3, TRI_3 5, QUAD_4 7."""
    w = AdfWriter()
    w._alloc(512)  # room for the root record at 266
    xy = np.asarray(xy, float)
    cx = w.add("CoordinateX", "DataArray_t", "R8", xy[:, 0].copy(), [len(xy)])
    cy = w.add("CoordinateY", "DataArray_t", "R8", xy[:, 1].copy(), [len(xy)])
    gc = w.add("GridCoordinates", "GridCoordinates_t", children=[("CoordinateX", cx), ("CoordinateY", cy)])
    zt = w.add("ZoneType", "ZoneType_t", "C1", b"Unstructured", [12])
    kids = [("ZoneType", zt), ("GridCoordinates", gc)]
    ncell = 0
    for name, typ, first, nodes in elements:
        conn = (np.asarray(nodes, np.int32) + 1).reshape(-1)
        er = w.add("ElementRange", "IndexRange_t", "I4", np.array([first, first + len(nodes) - 1], np.int32), [2])
        ec = w.add("ElementConnectivity", "DataArray_t", "I4", conn, [len(conn)])
        sec = w.add(name, "Elements_t", "I4", np.array([typ, 0], np.int32), [2],
                    children=[("ElementRange", er), ("ElementConnectivity", ec)])
        kids.append((name, sec))
        ncell += len(nodes) if typ != 3 else 0
    bck = []
    for name, ids in bcs:
        pl = w.add("PointList", "IndexArray_t", "I4", np.asarray(ids, np.int32), [1, len(ids)])
        gl = w.add("GridLocation", "GridLocation_t", "C1", b"EdgeCenter", [10])
        bck.append((name, w.add(name, "BC_t", "C1", b"BCGeneral", [9], children=[("PointList", pl), ("GridLocation", gl)])))
    kids.append(("ZoneBC", w.add("ZoneBC", "ZoneBC_t", children=bck)))
    zone = w.add("Zone", "Zone_t", "I4", np.array([len(xy), ncell, 0], np.int32), [1, 3], children=kids)
    base = w.add("Base", "CGNSBase_t", "I4", np.array([2, 2], np.int32), [2], children=[("Zone", zone)])
    ver = w.add("CGNSLibraryVersion", "CGNSLibraryVersion_t", "R4", np.array([3.1], np.float32), [1])
    w.finish(path, [("CGNSLibraryVersion", ver), ("Base", base)])


def read_cgns_py(path):
    """Independent (Python) ADF parse used to cross-check the C++ reader on real files."""
    d = open(path, "rb").read()
    ptr = lambda b: int(b[:8], 16) * 4096 + int(b[8:12], 16)

    def node(off):
        assert d[off:off + 4] == b"NoDe"
        nd = int(d[off + 128:off + 130], 16)
        return dict(name=d[off + 4:off + 36].decode().strip(), label=d[off + 36:off + 68].decode().strip(),
                    nsub=int(d[off + 68:off + 76], 16), sub=ptr(d[off + 84:off + 96]),
                    dtype=d[off + 96:off + 128].decode().strip(),
                    dims=[int(d[off + 130 + 8 * i:off + 138 + 8 * i], 16) for i in range(nd)], data=ptr(d[off + 230:off + 242]))

    def kids(n):
        return [node(ptr(d[n["sub"] + 16 + 44 * i + 32:n["sub"] + 16 + 44 * i + 44])) for i in range(n["nsub"])]

    def arr(n):
        cnt = int(np.prod(n["dims"]))
        dt = {"I4": np.int32, "R8": np.float64}[n["dtype"]]
        return np.frombuffer(d, dtype=dt, count=cnt, offset=n["data"] + 16).copy()

    root = node(d.find(b"NoDe"))
    base = [k for k in kids(root) if k["label"] == "CGNSBase_t"][0]
    zone = [k for k in kids(base) if k["label"] == "Zone_t"][0]
    zk = kids(zone)
    gc = [k for k in zk if k["label"] == "GridCoordinates_t"][0]
    co = {k["name"]: arr(k) for k in kids(gc)}
    xy = np.stack([co["CoordinateX"], co["CoordinateY"]], 1)
    elems = {}
    for sec in [k for k in zk if k["label"] == "Elements_t"]:
        typ = int(arr(sec)[0])
        sk = {k["name"]: k for k in kids(sec)}
        rng, conn = arr(sk["ElementRange"]), arr(sk["ElementConnectivity"])
        per = {3: 2, 5: 3, 7: 4}[typ]
        for i, eid in enumerate(range(rng[0], rng[1] + 1)):
            elems[eid] = conn[per * i:per * (i + 1)] - 1
    cptr, cind = [0], []
    for eid in sorted(elems):
        if len(elems[eid]) > 2:
            cind += list(elems[eid])
            cptr.append(len(cind))
    patches = []
    for zbc in [k for k in zk if k["label"] == "ZoneBC_t"]:
        for bc in kids(zbc):
            pl = [k for k in kids(bc) if k["name"] == "PointList"][0]
            pairs = []
            for eid in arr(pl):
                pairs += list(elems[int(eid)])
            patches.append((bc["name"], pairs))
    return xy, np.array(cptr, np.int32), np.array(cind, np.int32), patches
