// partition.cu -- the reference's two METIS entry points and the per-partition grid-file layout of its
// PhasePartitionGrid utility (host integer work).
//
//   FiniteVolumeGrid2D::partition        UG/FiniteVolumeGrid2D.cpp:287-297  METIS_PartMeshDual(ne, nn, eptr, eind, ncommon = 2)
//   PhasePartitionGrid                   U/utilities/PhasePartitionGrid.cpp:42-50   METIS_PartGraphRecursive on the
//                                        face-neighbour graph (connectivityGraph, UG/FiniteVolumeGrid2D.cpp:243-254)
//   PhasePartitionGrid, per partition    :56-127  cell order (owned ascending, then halo cells in discovery order),
//                                        ProcNo, node renumbering in first-use order, 1-based element lists, patches
//                                        as node pairs; :150-153 the GlobalID / ProcNo fields
// METIS itself is the libmetis_static.a that ships with the CUDA toolkit (for cuSOLVER; 64-bit idx_t), linked when
// the build finds it (PHB_HAVE_METIS).  The reference passes 32-bit ints to a system METIS: the partition VECTOR is
// therefore an input of every parity check, never a parity target.
#include <algorithm>
#include <unordered_map>
#include <unordered_set>

#include "structs.cuh"

using namespace phb;

#ifdef PHB_HAVE_METIS
typedef long long metis_idx_t;
typedef float metis_real_t;
extern "C" int METIS_PartGraphRecursive(metis_idx_t *, metis_idx_t *, metis_idx_t *, metis_idx_t *, metis_idx_t *, metis_idx_t *,
                                        metis_idx_t *, metis_idx_t *, metis_real_t *, metis_real_t *, metis_idx_t *, metis_idx_t *,
                                        metis_idx_t *);
extern "C" int METIS_PartMeshDual(metis_idx_t *, metis_idx_t *, metis_idx_t *, metis_idx_t *, metis_idx_t *, metis_idx_t *,
                                  metis_idx_t *, metis_idx_t *, metis_real_t *, metis_idx_t *, metis_idx_t *, metis_idx_t *,
                                  metis_idx_t *);
#endif

extern "C" {

// method 0: METIS_PartMeshDual, ncommon 2 (the solver's own partition); 1: METIS_PartGraphRecursive on the
// face-neighbour graph (the PhasePartitionGrid utility).  objective = edge cut reported by METIS.
int phb_partition_metis(const phb_mesh *g, int nParts, int method, int *cellPartition, long long *objective) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(g && cellPartition && nParts >= 1 && (method == 0 || method == 1), "phb_partition_metis: bad argument");
  PHB_REQUIRE(g->finalized, "phb_partition_metis: mesh must be finalized");
#ifdef PHB_HAVE_METIS
  const int N = g->nCells;
  if (nParts == 1) { std::fill(cellPartition, cellPartition + N, 0); if (objective) *objective = 0; return PHB_OK; }
  metis_idx_t ne = N, nn = g->nNodes, ncommon = 2, nparts = nParts, objval = 0, ncon = 1;
  std::vector<metis_idx_t> epart(N), npart(g->nNodes);
  int rc;
  if (method == 0) {
    std::vector<metis_idx_t> eptr(g->cptr.begin(), g->cptr.end()), eind(g->cind.begin(), g->cind.end());
    rc = METIS_PartMeshDual(&ne, &nn, eptr.data(), eind.data(), nullptr, nullptr, &ncommon, &nparts, nullptr, nullptr, &objval,
                            epart.data(), npart.data());
  } else {
    std::vector<metis_idx_t> xadj(g->ilPtr.begin(), g->ilPtr.end()), adj(g->ilCell.begin(), g->ilCell.end());
    rc = METIS_PartGraphRecursive(&ne, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &nparts, nullptr, nullptr, nullptr,
                                  &objval, epart.data());
  }
  if (rc != 1) { set_error("phb_partition_metis: METIS returned %d", rc); return PHB_ERR_STATE; }
  for (int i = 0; i < N; ++i) cellPartition[i] = (int)epart[i];
  if (objective) *objective = (long long)objval;
  return PHB_OK;
#else
  (void)objective;
  set_error("phb_partition_metis: built without METIS (libmetis_static.a of the CUDA toolkit not found)");
  return PHB_ERR_UNSUPPORTED;
#endif
  PHB_TRY_END
}

struct phb_partfile {
  std::vector<int> localCells, owningProc, eptr, eind;   // GlobalID, ProcNo, element lists (eind 1-based local node ids)
  std::vector<double> nodes;                              // x, y interleaved, local node order
  std::vector<std::string> patchNames;
  std::vector<std::vector<int>> patchNodes;               // node pairs, 1-based local ids
};

// The content PhasePartitionGrid writes into solution/Proc<proc>/Grid.cgns (:56-153)
int phb_partition_file_build(const phb_mesh *g, const int *part, int proc, double minBufferWidth, phb_partfile **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(g && part && out && proc >= 0, "phb_partition_file_build: bad argument");
  PHB_REQUIRE(g->finalized, "phb_partition_file_build: mesh must be finalized");
  const int N = g->nCells;
  std::unique_ptr<phb_partfile> F(new phb_partfile());
  // cellLinks = interior links then diagonal links, in push order (UG/Cell/Cell.cpp:74-94)
  auto for_links = [&](int c, auto f) {
    for (int j = g->ilPtr[c]; j < g->ilPtr[c + 1]; ++j) f(g->ilCell[j]);
    for (int j = g->dlPtr[c]; j < g->dlPtr[c + 1]; ++j) f(g->dlCell[j]);
  };
  std::vector<int> boundaryCells;
  for (int c = 0; c < N; ++c)
    if (part[c] == proc) {
      F->localCells.push_back(c);
      F->owningProc.push_back(proc);
      bool touches = false;
      for_links(c, [&](int nb) { if (part[nb] != proc) touches = true; });
      if (touches) boundaryCells.push_back(c);
    }
  std::unordered_set<int> seen;
  for (int c : boundaryCells) {
    for_links(c, [&](int nb) {
      if (part[nb] != proc && seen.insert(nb).second) { F->localCells.push_back(nb); F->owningProc.push_back(part[nb]); }
    });
    if (minBufferWidth > 0.)   // itemsCoveredBy(Circle(centroid, width)): |c_k - c| <= width, in cell id order
      for (int k = 0; k < N; ++k) {
        const double dx = g->cCx[k] - g->cCx[c], dy = g->cCy[k] - g->cCy[c];
        if (dx * dx + dy * dy <= minBufferWidth * minBufferWidth && part[k] != proc && seen.insert(k).second) {
          F->localCells.push_back(k);
          F->owningProc.push_back(part[k]);
        }
      }
  }
  std::unordered_map<int, int> g2l;
  F->eptr.push_back(0);
  for (int c : F->localCells) {
    F->eptr.push_back(F->eptr.back() + (g->cptr[c + 1] - g->cptr[c]));
    for (int j = g->cptr[c]; j < g->cptr[c + 1]; ++j) {
      const int nd = g->cind[j];
      auto ins = g2l.insert({nd, (int)(F->nodes.size() / 2)});
      if (ins.second) { F->nodes.push_back(g->nodeX[nd]); F->nodes.push_back(g->nodeY[nd]); }
      F->eind.push_back(ins.first->second + 1);
    }
  }
  for (size_t p = 0; p < g->patchNames.size(); ++p) {
    std::vector<int> pairs;
    for (int f = 0; f < g->nFaces; ++f) {
      if (g->fR[f] >= 0 || g->fPatch[f] != (int)p) continue;
      auto a = g2l.find(g->fN1[f]), b = g2l.find(g->fN2[f]);
      if (a != g2l.end() && b != g2l.end()) { pairs.push_back(a->second + 1); pairs.push_back(b->second + 1); }
    }
    if (!pairs.empty()) { F->patchNames.push_back(g->patchNames[p]); F->patchNodes.push_back(pairs); }
  }
  *out = F.release();
  return PHB_OK;
  PHB_TRY_END
}

// sizes: [nCells, nNodes, len(eind), nPatches]
int phb_partition_file_sizes(const phb_partfile *f, long long out[4]) {
  PHB_REQUIRE(f && out, "phb_partition_file_sizes: NULL argument");
  out[0] = (long long)f->localCells.size(); out[1] = (long long)f->nodes.size() / 2;
  out[2] = (long long)f->eind.size(); out[3] = (long long)f->patchNames.size();
  return PHB_OK;
}
int phb_partition_file_get(const phb_partfile *f, int *globalId, int *procNo, double *nodesXY, int *eptr, int *eind) {
  PHB_REQUIRE(f, "phb_partition_file_get: NULL argument");
  if (globalId) std::copy(f->localCells.begin(), f->localCells.end(), globalId);
  if (procNo) std::copy(f->owningProc.begin(), f->owningProc.end(), procNo);
  if (nodesXY) std::copy(f->nodes.begin(), f->nodes.end(), nodesXY);
  if (eptr) std::copy(f->eptr.begin(), f->eptr.end(), eptr);
  if (eind) std::copy(f->eind.begin(), f->eind.end(), eind);
  return PHB_OK;
}
// patch p: name (cap bytes) and node pairs; returns the number of node ids (2 per face) or < 0
long long phb_partition_file_patch(const phb_partfile *f, int p, char *name, int cap, int *nodePairs) {
  if (!f || p < 0 || p >= (int)f->patchNames.size()) return PHB_ERR_ARG;
  if (name && cap > 0) { strncpy(name, f->patchNames[p].c_str(), cap - 1); name[cap - 1] = 0; }
  if (nodePairs) std::copy(f->patchNodes[p].begin(), f->patchNodes[p].end(), nodePairs);
  return (long long)f->patchNodes[p].size();
}
int phb_partition_file_destroy(phb_partfile *f) { delete f; return PHB_OK; }

}  // extern "C"
