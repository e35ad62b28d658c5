// ingest.cu -- mesh ingest on the wire formats the reference reads:
//   * classic ADF-format CGNS files (all three shipped .cgns meshes are ADF, not HDF5),
//     with the semantics of CgnsUnstructuredGrid::load (UG/CgnsUnstructuredGrid.cpp:13-105,
//     S/CgnsFile.cpp:183-241): one 2-D base, one Unstructured zone; every Elements_t section
//     is merged into one table indexed by ELEMENT ID; connectivity becomes 0-based; every
//     BC_t PointList / PointRange entry is an element id whose two nodes define a patch
//     face; cells are the elements with more than two nodes, in element-id order.
//   * uniform ("red") refinement of triangle / quad meshes, used to scale the 15 316-triangle
//     cylinder mesh to the 16M-cell configuration.
// No cgnslib / HDF5: the ADF container is parsed directly (node = 246-byte "NoDe" record,
// sub-node table "SNTb", data chunk "DaTa"; file offset = block * 4096 + offset).
#include <algorithm>
#include <fstream>
#include <map>

#include "structs.cuh"

namespace {

struct AdfNode {
  std::string name, label, dtype;
  std::vector<long long> dims;
  long long nSub = 0;
  size_t sub = 0, data = 0;
  int nChunks = 0;
};

struct AdfFile {
  std::vector<unsigned char> d;
  long long hex(size_t off, int n) const {
    long long v = 0;
    for (int i = 0; i < n; ++i) {
      const unsigned char c = d.at(off + i);
      int x = c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : -1;
      if (x < 0) throw std::runtime_error("ADF: bad hex digit");
      v = v * 16 + x;
    }
    return v;
  }
  size_t ptr(size_t off) const { return (size_t)hex(off, 8) * 4096 + (size_t)hex(off + 8, 4); }
  static std::string trim(const std::string &s) {
    size_t e = s.find_last_not_of(" \0", std::string::npos, 2);
    return e == std::string::npos ? "" : s.substr(0, e + 1);
  }
  std::string str(size_t off, int n) const { return trim(std::string((const char *)&d.at(off), n)); }
  AdfNode node(size_t off) const {
    if (off + 246 > d.size() || memcmp(&d[off], "NoDe", 4) != 0) throw std::runtime_error("ADF: node tag not found");
    AdfNode n;
    n.name = str(off + 4, 32); n.label = str(off + 36, 32);
    n.nSub = hex(off + 68, 8); n.sub = ptr(off + 84);
    n.dtype = str(off + 96, 32);
    const int nd = (int)hex(off + 128, 2);
    for (int i = 0; i < nd && i < 12; ++i) n.dims.push_back(hex(off + 130 + 8 * i, 8));
    n.nChunks = (int)hex(off + 226, 4);
    n.data = ptr(off + 230);
    return n;
  }
  std::vector<AdfNode> children(const AdfNode &n) const {
    std::vector<AdfNode> out;
    if (n.nSub == 0) return out;
    if (memcmp(&d.at(n.sub), "SNTb", 4) != 0) throw std::runtime_error("ADF: sub-node table tag not found");
    size_t p = n.sub + 16;
    for (long long i = 0; i < n.nSub; ++i, p += 44) out.push_back(node(ptr(p + 32)));
    return out;
  }
  size_t count(const AdfNode &n) const {
    size_t c = 1;
    for (long long v : n.dims) c *= (size_t)v;
    return n.dims.empty() ? 0 : c;
  }
  const unsigned char *payload(const AdfNode &n) const {
    if (n.nChunks != 1) throw std::runtime_error("ADF: node \"" + n.name + "\" has " + std::to_string(n.nChunks) + " data chunks");
    if (memcmp(&d.at(n.data), "DaTa", 4) != 0) throw std::runtime_error("ADF: data chunk tag not found");
    return &d.at(n.data + 16);
  }
  std::vector<int> i4(const AdfNode &n) const {
    if (n.dtype != "I4") throw std::runtime_error("ADF: node \"" + n.name + "\" is not I4");
    std::vector<int> v(count(n));
    if (!v.empty()) memcpy(v.data(), payload(n), v.size() * 4);
    return v;
  }
  std::vector<double> r8(const AdfNode &n) const {
    if (n.dtype != "R8") throw std::runtime_error("ADF: node \"" + n.name + "\" is not R8");
    std::vector<double> v(count(n));
    if (!v.empty()) memcpy(v.data(), payload(n), v.size() * 8);
    return v;
  }
  std::string c1(const AdfNode &n) const {
    const size_t c = count(n);
    return c ? trim(std::string((const char *)payload(n), c)) : std::string();
  }
};

const AdfNode *find(const std::vector<AdfNode> &v, const std::string &label, const std::string &name = "") {
  for (const AdfNode &n : v)
    if (n.label == label && (name.empty() || n.name == name)) return &n;
  return nullptr;
}

int nodes_per_element(int type) {
  switch (type) {
    case 3: return 2;   // BAR_2
    case 5: return 3;   // TRI_3
    case 7: return 4;   // QUAD_4
    default: return -1;
  }
}

}  // namespace

extern "C" {

int phb_mesh_read_cgns(phb_ctx *ctx, const char *filename, phb_mesh **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(ctx && filename && out, "phb_mesh_read_cgns: NULL argument");
  AdfFile f;
  {
    std::ifstream in(filename, std::ios::binary);
    PHB_REQUIRE((bool)in, "phb_mesh_read_cgns: cannot open \"%s\"", filename);
    f.d.assign(std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>());
  }
  PHB_REQUIRE(f.d.size() > 512 && memcmp(&f.d[4], "ADF Database Version", 20) == 0,
              "phb_mesh_read_cgns: \"%s\" is not an ADF-format CGNS file (HDF5 files are not supported)", filename);
  size_t rootOff = 0;
  for (size_t i = 0; i + 4 < std::min<size_t>(f.d.size(), 8192); ++i)
    if (memcmp(&f.d[i], "NoDe", 4) == 0) { rootOff = i; break; }
  PHB_REQUIRE(rootOff, "phb_mesh_read_cgns: root node not found");
  const AdfNode root = f.node(rootOff);
  // base with cellDim == 2 (CgnsUnstructuredGrid.cpp:22-28)
  const AdfNode *base = nullptr;
  const std::vector<AdfNode> top = f.children(root);
  for (const AdfNode &n : top)
    if (n.label == "CGNSBase_t" && f.i4(n).at(0) == 2) { base = &n; break; }
  PHB_REQUIRE(base, "phb_mesh_read_cgns: no base with cell dimension 2");
  const std::vector<AdfNode> zones = f.children(*base);
  const AdfNode *zone = find(zones, "Zone_t");
  PHB_REQUIRE(zone, "phb_mesh_read_cgns: no zone");
  const std::vector<AdfNode> zc = f.children(*zone);
  const AdfNode *zt = find(zc, "ZoneType_t");
  PHB_REQUIRE(zt && f.c1(*zt) == "Unstructured", "phb_mesh_read_cgns: zone is not Unstructured");
  const AdfNode *gc = find(zc, "GridCoordinates_t");
  PHB_REQUIRE(gc, "phb_mesh_read_cgns: no GridCoordinates");
  const std::vector<AdfNode> coords = f.children(*gc);
  const AdfNode *cx = find(coords, "DataArray_t", "CoordinateX"), *cy = find(coords, "DataArray_t", "CoordinateY");
  PHB_REQUIRE(cx && cy, "phb_mesh_read_cgns: CoordinateX/Y missing");
  const std::vector<double> X = f.r8(*cx), Y = f.r8(*cy);
  const int nNodes = (int)X.size();
  // element table indexed by element id (1-based in the file)
  std::map<int, std::vector<int>> elems;
  for (const AdfNode &sec : zc) {
    if (sec.label != "Elements_t") continue;
    const int type = f.i4(sec).at(0);
    const std::vector<AdfNode> sc = f.children(sec);
    const AdfNode *er = find(sc, "IndexRange_t", "ElementRange"), *ec = find(sc, "DataArray_t", "ElementConnectivity");
    PHB_REQUIRE(er && ec, "phb_mesh_read_cgns: section \"%s\" lacks range/connectivity", sec.name.c_str());
    const std::vector<int> range = f.i4(*er), conn = f.i4(*ec);
    size_t p = 0;
    for (int id = range.at(0); id <= range.at(1); ++id) {
      int t = type;
      if (type == 20) t = conn.at(p++);  // MIXED: the element type precedes its nodes
      const int k = nodes_per_element(t);
      PHB_REQUIRE(k > 0, "phb_mesh_read_cgns: unsupported element type %d (BAR_2, TRI_3, QUAD_4, MIXED only)", t);
      std::vector<int> nd(k);
      for (int q = 0; q < k; ++q) nd[q] = conn.at(p++) - 1;
      elems[id] = nd;
    }
  }
  std::vector<double> xy(2 * (size_t)nNodes);
  for (int i = 0; i < nNodes; ++i) { xy[2 * i] = X[i]; xy[2 * i + 1] = Y[i]; }
  std::vector<int> cptr(1, 0), cind;
  for (const auto &e : elems)
    if (e.second.size() > 2) {
      cind.insert(cind.end(), e.second.begin(), e.second.end());
      cptr.push_back((int)cind.size());
    }
  PHB_REQUIRE(cptr.size() > 1, "phb_mesh_read_cgns: no 2-D elements");
  phb_mesh *m = nullptr;
  PHB_CHECK(phb_mesh_create(ctx, nNodes, xy.data(), (int)cptr.size() - 1, cptr.data(), cind.data(), &m));
  std::unique_ptr<phb_mesh> guard(m);
  const AdfNode *zbc = find(zc, "ZoneBC_t");
  if (zbc)
    for (const AdfNode &bc : f.children(*zbc)) {
      if (bc.label != "BC_t") continue;
      const std::vector<AdfNode> bcc = f.children(bc);
      std::vector<int> ids;
      if (const AdfNode *pl = find(bcc, "IndexArray_t", "PointList")) ids = f.i4(*pl);
      else if (const AdfNode *pr = find(bcc, "IndexRange_t", "PointRange")) {
        const std::vector<int> r = f.i4(*pr);
        for (int id = r.at(0); id <= r.at(1); ++id) ids.push_back(id);
      }
      std::vector<int> pairs;
      for (int id : ids) {
        auto it = elems.find(id);
        PHB_REQUIRE(it != elems.end() && it->second.size() == 2, "phb_mesh_read_cgns: BC \"%s\" names element %d, not a BAR_2",
                    bc.name.c_str(), id);
        pairs.push_back(it->second[0]); pairs.push_back(it->second[1]);
      }
      if (phb_mesh_add_patch_by_nodes(m, bc.name.c_str(), (int)pairs.size() / 2, pairs.data()) < 0) return PHB_ERR_ARG;
    }
  *out = guard.release();
  return PHB_OK;
  PHB_TRY_END
}

// Uniform refinement: every edge is split at its midpoint; a triangle becomes 4 triangles, a quad 4 quads
// (centre node added).  Children keep the parent's orientation and are numbered parent-major; the two halves
// of a patch face stay in the patch.  `levels` rounds are applied; the input may be finalized or not.
int phb_mesh_refine(phb_ctx *ctx, const phb_mesh *in, int levels, phb_mesh **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(ctx && in && out && levels >= 1, "phb_mesh_refine: bad argument");
  std::vector<double> X = in->nodeX, Y = in->nodeY;
  std::vector<int> cptr = in->cptr, cind = in->cind;
  // patch faces as node pairs
  std::vector<std::vector<int>> patchPairs(in->patchNames.size());
  for (int f = 0; f < in->nFaces; ++f)
    if (in->fPatch[f] >= 0) { patchPairs[in->fPatch[f]].push_back(in->fN1[f]); patchPairs[in->fPatch[f]].push_back(in->fN2[f]); }
  for (int lv = 0; lv < levels; ++lv) {
    std::map<std::pair<int, int>, int> mid;
    auto midpoint = [&](int a, int b) {
      const std::pair<int, int> key(std::min(a, b), std::max(a, b));
      auto it = mid.find(key);
      if (it != mid.end()) return it->second;
      const int id = (int)X.size();
      X.push_back(0.5 * (X[a] + X[b])); Y.push_back(0.5 * (Y[a] + Y[b]));
      mid[key] = id;
      return id;
    };
    std::vector<int> np(1, 0), ni;
    const int nc = (int)cptr.size() - 1;
    for (int c = 0; c < nc; ++c) {
      const int *v = &cind[cptr[c]];
      const int k = cptr[c + 1] - cptr[c];
      if (k == 3) {
        const int a = midpoint(v[0], v[1]), b = midpoint(v[1], v[2]), d = midpoint(v[2], v[0]);
        const int t[12] = {v[0], a, d, a, v[1], b, d, b, v[2], a, b, d};
        for (int q = 0; q < 4; ++q) { ni.insert(ni.end(), t + 3 * q, t + 3 * q + 3); np.push_back((int)ni.size()); }
      } else if (k == 4) {
        const int a = midpoint(v[0], v[1]), b = midpoint(v[1], v[2]), e = midpoint(v[2], v[3]), d = midpoint(v[3], v[0]);
        const int ctr = (int)X.size();
        X.push_back(0.25 * (X[v[0]] + X[v[1]] + X[v[2]] + X[v[3]])); Y.push_back(0.25 * (Y[v[0]] + Y[v[1]] + Y[v[2]] + Y[v[3]]));
        const int t[16] = {v[0], a, ctr, d, a, v[1], b, ctr, ctr, b, v[2], e, d, ctr, e, v[3]};
        for (int q = 0; q < 4; ++q) { ni.insert(ni.end(), t + 4 * q, t + 4 * q + 4); np.push_back((int)ni.size()); }
      } else {
        PHB_REQUIRE(false, "phb_mesh_refine: only triangles and quads can be refined (cell %d has %d nodes)", c, k);
      }
    }
    for (auto &pp : patchPairs) {
      std::vector<int> q;
      for (size_t i = 0; i + 1 < pp.size(); i += 2) {
        const int mm = midpoint(pp[i], pp[i + 1]);
        q.push_back(pp[i]); q.push_back(mm); q.push_back(mm); q.push_back(pp[i + 1]);
      }
      pp.swap(q);
    }
    cptr.swap(np); cind.swap(ni);
  }
  std::vector<double> xy(2 * X.size());
  for (size_t i = 0; i < X.size(); ++i) { xy[2 * i] = X[i]; xy[2 * i + 1] = Y[i]; }
  phb_mesh *m = nullptr;
  PHB_CHECK(phb_mesh_create(ctx, (int)X.size(), xy.data(), (int)cptr.size() - 1, cptr.data(), cind.data(), &m));
  std::unique_ptr<phb_mesh> guard(m);
  for (size_t p = 0; p < patchPairs.size(); ++p)
    if (phb_mesh_add_patch_by_nodes(m, in->patchNames[p].c_str(), (int)patchPairs[p].size() / 2, patchPairs[p].data()) < 0)
      return PHB_ERR_ARG;
  *out = guard.release();
  return PHB_OK;
  PHB_TRY_END
}

int phb_mesh_patch_name(const phb_mesh *m, int id, char *out, int cap) {
  PHB_REQUIRE(m && out && cap > 0, "phb_mesh_patch_name: bad argument");
  PHB_REQUIRE(id >= 0 && id < (int)m->patchNames.size(), "phb_mesh_patch_name: id %d out of range", id);
  snprintf(out, cap, "%s", m->patchNames[id].c_str());
  return PHB_OK;
}

}  // extern "C"
