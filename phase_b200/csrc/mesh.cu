// mesh.cu -- host-side mesh build (integer artefacts I1-I5 must be bit-exact with
// the reference's numbering), geometry factors (G1-G4), canonical CSR pattern
// with the face->slot map, and the SoA upload the kernels read.
//
// Reference semantics reproduced (file:line under /root/reference/src/2D/Unstructured):
//   FiniteVolumeGrid2D/FiniteVolumeGrid2D.cpp:20-35,84-114  node/cell/face numbering
//   FiniteVolumeGrid2D/FiniteVolumeGrid2D.cpp:396-450       link order (ascending face id), diagonal links
//   FiniteVolumeGrid2D/Face/Face.cpp:9-18,48-64             face centroid/normal/outwardNorm/weights
//   FiniteVolumeGrid2D/Cell/Cell.cpp:8-29                   cell area + centroid
//   FiniteVolumeGrid2D/StructuredRectilinearGrid.cpp:40-95,175-194
//   FiniteVolumeGrid2D/FiniteVolumeGrid2D.cpp:276-392,460-511  partition, buffer/send groups
//   FiniteVolume/Equation/IndexMap.cpp:15-40                local/global row numbering
#include <algorithm>
#include <cmath>
#include <numeric>

#include "structs.cuh"

using phb::set_error;

namespace {

inline uint64_t pair_key(int a, int b) {
  uint32_t lo = (uint32_t)std::min(a, b), hi = (uint32_t)std::max(a, b);
  return ((uint64_t)lo << 32) | hi;
}
inline size_t mix(uint64_t k) {
  k ^= k >> 31;
  k *= 0x7fb5d329728ea185ULL;
  k ^= k >> 27;
  k *= 0x81dadef4bc2dd44dULL;
  k ^= k >> 33;
  return (size_t)k;
}
int dir_find(const phb_mesh *m, int a, int b) {
  const size_t mask = m->hVal.size() - 1;
  const uint64_t key = pair_key(a, b);
  for (size_t i = mix(key) & mask;; i = (i + 1) & mask) {
    if (m->hVal[i] < 0) return -1;
    if (m->hKey[i] == key) return m->hVal[i];
  }
}
void dir_put(phb_mesh *m, int a, int b, int id) {
  const size_t mask = m->hVal.size() - 1;
  const uint64_t key = pair_key(a, b);
  size_t i = mix(key) & mask;
  while (m->hVal[i] >= 0) i = (i + 1) & mask;
  m->hKey[i] = key;
  m->hVal[i] = id;
}

// I1: faces are numbered in order of first appearance while walking each cell's
// edges (n_k, n_k+1); the first cell to touch a face is its lCell.
int build_faces(phb_mesh *m) {
  const int N = m->nCells;
  const size_t nInd = m->cind.size();
  size_t cap = 16;
  while (cap < 2 * nInd + 16) cap <<= 1;
  m->hKey.assign(cap, 0);
  m->hVal.assign(cap, -1);
  m->fN1.clear(); m->fN2.clear(); m->fL.clear(); m->fR.clear();
  m->fN1.reserve(nInd / 2 + 16); m->fN2.reserve(nInd / 2 + 16);
  m->fL.reserve(nInd / 2 + 16); m->fR.reserve(nInd / 2 + 16);
  m->vol.resize(N); m->cCx.resize(N); m->cCy.resize(N);
  for (int c = 0; c < N; ++c) {
    const int b = m->cptr[c], nv = m->cptr[c + 1] - b;
    if (nv < 3) {
      set_error("mesh: cell %d has %d nodes", c, nv);
      return PHB_ERR_ARG;
    }
    // polygon area and centroid about the first vertex
    const double x0 = m->nodeX[m->cind[b]], y0 = m->nodeY[m->cind[b]];
    double a2 = 0., sx = 0., sy = 0.;
    for (int k = 0; k < nv; ++k) {
      const int n1 = m->cind[b + k], n2 = m->cind[b + (k + 1) % nv];
      const double x1 = m->nodeX[n1] - x0, y1 = m->nodeY[n1] - y0;
      const double x2 = m->nodeX[n2] - x0, y2 = m->nodeY[n2] - y0;
      const double ai = x1 * y2 - x2 * y1;
      a2 += ai;
      sx += ai * (x1 + x2);
      sy += ai * (y1 + y2);
    }
    m->vol[c] = std::fabs(0.5 * a2);
    m->cCx[c] = sx / (3. * a2) + x0;
    m->cCy[c] = sy / (3. * a2) + y0;
    for (int k = 0; k < nv; ++k) {
      const int n1 = m->cind[b + k], n2 = m->cind[b + (k + 1) % nv];
      int f = dir_find(m, n1, n2);
      if (f < 0) {
        f = (int)m->fL.size();
        m->fN1.push_back(n1);
        m->fN2.push_back(n2);
        m->fL.push_back(c);
        m->fR.push_back(-1);
        dir_put(m, n1, n2, f);
      } else {
        if (m->fR[f] >= 0) {
          set_error("mesh: face (%d,%d) shared by more than two cells", n1, n2);
          return PHB_ERR_ARG;
        }
        m->fR[f] = c;
      }
    }
  }
  m->nFaces = (int)m->fL.size();
  m->fPatch.assign(m->nFaces, -1);
  return PHB_OK;
}

// G1, G3, G4 per FACE: both directed links of a face share |S_f|, g_f and r/|r|^2
// up to sign, so they are stored once.
int build_geometry(phb_mesh *m) {
  const int F = m->nFaces;
  m->fCx.resize(F); m->fCy.resize(F); m->fSx.resize(F); m->fSy.resize(F);
  m->fG.resize(F); m->fW.resize(F); m->fQx.resize(F); m->fQy.resize(F);
  for (int f = 0; f < F; ++f) {
    const double lx = m->nodeX[m->fN1[f]], ly = m->nodeY[m->fN1[f]];
    const double rx = m->nodeX[m->fN2[f]], ry = m->nodeY[m->fN2[f]];
    const double cx = 0.5 * (lx + rx), cy = 0.5 * (ly + ry);
    double nx = ry - ly, ny = -(rx - lx);  // normalVec of the tangent
    const int l = m->fL[f], r = m->fR[f];
    const double dl = (cx - m->cCx[l]) * nx + (cy - m->cCy[l]) * ny;
    if (!(dl > 0.)) { nx = -nx; ny = -ny; }
    m->fCx[f] = cx; m->fCy[f] = cy; m->fSx[f] = nx; m->fSy[f] = ny;
    double qx, qy;
    if (r >= 0) {
      const double dr = (cx - m->cCx[r]) * nx + (cy - m->cCy[r]) * ny;
      if (dr > 0.) {
        set_error("mesh: face %d is not outward for both of its cells (non-convex link)", f);
        return PHB_ERR_UNSUPPORTED;
      }
      qx = m->cCx[r] - m->cCx[l]; qy = m->cCy[r] - m->cCy[l];
      const double l1 = std::hypot(cx - m->cCx[r], cy - m->cCy[r]);
      const double l2 = std::hypot(cx - m->cCx[l], cy - m->cCy[l]);
      m->fW[f] = l1 / (l1 + l2);
    } else {
      qx = cx - m->cCx[l]; qy = cy - m->cCy[l];
      m->fW[f] = 1.;
    }
    const double mm = qx * qx + qy * qy;
    m->fG[f] = (qx * nx + qy * ny) / mm;
    m->fQx[f] = qx / mm;
    m->fQy[f] = qy / mm;
  }
  return PHB_OK;
}

// I2: per-cell links in ascending face id; diagonal links appended per node walk.
void build_links(phb_mesh *m) {
  const int N = m->nCells, F = m->nFaces;
  m->ilPtr.assign(N + 1, 0);
  m->blPtr.assign(N + 1, 0);
  for (int f = 0; f < F; ++f) {
    if (m->fR[f] < 0) m->blPtr[m->fL[f] + 1]++;
    else { m->ilPtr[m->fL[f] + 1]++; m->ilPtr[m->fR[f] + 1]++; }
  }
  std::partial_sum(m->ilPtr.begin(), m->ilPtr.end(), m->ilPtr.begin());
  std::partial_sum(m->blPtr.begin(), m->blPtr.end(), m->blPtr.begin());
  m->ilFace.resize(m->ilPtr[N]); m->ilCell.resize(m->ilPtr[N]); m->blFace.resize(m->blPtr[N]);
  std::vector<int> ic(N, 0), bc(N, 0);
  for (int f = 0; f < F; ++f) {
    const int l = m->fL[f], r = m->fR[f];
    if (r < 0) m->blFace[m->blPtr[l] + bc[l]++] = f;
    else {
      int j = m->ilPtr[l] + ic[l]++;
      m->ilFace[j] = f; m->ilCell[j] = r;
      j = m->ilPtr[r] + ic[r]++;
      m->ilFace[j] = f; m->ilCell[j] = l;
    }
  }
  // node -> cells (ascending cell id)
  std::vector<int> npPtr(m->nNodes + 1, 0);
  for (int n : m->cind) npPtr[n + 1]++;
  std::partial_sum(npPtr.begin(), npPtr.end(), npPtr.begin());
  std::vector<int> npCell(m->cind.size()), fill(m->nNodes, 0);
  for (int c = 0; c < N; ++c)
    for (int k = m->cptr[c]; k < m->cptr[c + 1]; ++k) {
      const int n = m->cind[k];
      npCell[npPtr[n] + fill[n]++] = c;
    }
  m->dlPtr.assign(N + 1, 0);
  m->dlCell.clear();
  for (int c = 0; c < N; ++c) {
    for (int k = m->cptr[c]; k < m->cptr[c + 1]; ++k) {
      const int n = m->cind[k];
      for (int q = npPtr[n]; q < npPtr[n + 1]; ++q) {
        const int kc = npCell[q];
        if (kc == c) continue;
        bool share = false;
        for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) share |= m->ilCell[j] == kc;
        if (!share) m->dlCell.push_back(kc);
      }
    }
    m->dlPtr[c + 1] = (int)m->dlCell.size();
  }
}

void default_parallel(phb_mesh *m) {
  const int N = m->nCells;
  m->rank = 0; m->nProcs = 1; m->nLocal = N; m->rowOffset = 0;
  m->owner.assign(N, 0);
  m->globalId.resize(N); m->localRow.resize(N); m->globalRow.resize(N);
  std::iota(m->globalId.begin(), m->globalId.end(), 0);
  std::iota(m->localRow.begin(), m->localRow.end(), 0);
  std::iota(m->globalRow.begin(), m->globalRow.end(), 0);
  m->bufPtr.assign(2, 0); m->sendPtr.assign(2, 0);
  m->bufCell.clear(); m->sendCell.clear();
}

int mesh_from_arrays(phb_ctx *ctx, int nNodes, const double *xy, int nCells, const int *cptr,
                     const int *cind, phb_mesh **out) {
  PHB_REQUIRE(ctx && xy && cptr && cind && out, "phb_mesh_create: NULL argument");
  PHB_REQUIRE(nNodes > 0 && nCells > 0, "phb_mesh_create: empty mesh");
  std::unique_ptr<phb_mesh> m(new phb_mesh());
  m->ctx = ctx;
  m->nNodes = nNodes; m->nCells = nCells;
  m->nodeX.resize(nNodes); m->nodeY.resize(nNodes);
  for (int i = 0; i < nNodes; ++i) { m->nodeX[i] = xy[2 * i]; m->nodeY[i] = xy[2 * i + 1]; }
  m->cptr.assign(cptr, cptr + nCells + 1);
  PHB_REQUIRE(cptr[0] == 0, "phb_mesh_create: cptr[0] != 0");
  m->cind.assign(cind, cind + cptr[nCells]);
  for (int n : m->cind) PHB_REQUIRE(n >= 0 && n < nNodes, "phb_mesh_create: node id %d out of range", n);
  PHB_CHECK(build_faces(m.get()));
  default_parallel(m.get());
  *out = m.release();
  return PHB_OK;
}

void rect_nodes(int nx, int ny, double w, double h, std::vector<double> &xy) {
  const double hx0 = w / nx, hy0 = h / ny;
  xy.resize(2 * (size_t)(nx + 1) * (ny + 1));
  for (int j = 0; j <= ny; ++j)
    for (int i = 0; i <= nx; ++i) {
      const size_t k = (size_t)j * (nx + 1) + i;
      xy[2 * k] = i * hx0;
      xy[2 * k + 1] = j * hy0;
    }
}
int rect_patches(phb_mesh *m, int nx, int ny) {
  const int nnx = nx + 1;
  std::vector<int> pr(2 * (size_t)std::max(nx, ny));
  for (int j = 0; j < ny; ++j) { pr[2 * j] = j * nnx; pr[2 * j + 1] = (j + 1) * nnx; }
  if (phb_mesh_add_patch_by_nodes(m, "x-", ny, pr.data()) < 0) return PHB_ERR_ARG;
  for (int j = 0; j < ny; ++j) { pr[2 * j] = j * nnx + nx; pr[2 * j + 1] = (j + 1) * nnx + nx; }
  if (phb_mesh_add_patch_by_nodes(m, "x+", ny, pr.data()) < 0) return PHB_ERR_ARG;
  for (int i = 0; i < nx; ++i) { pr[2 * i] = i; pr[2 * i + 1] = i + 1; }
  if (phb_mesh_add_patch_by_nodes(m, "y-", nx, pr.data()) < 0) return PHB_ERR_ARG;
  for (int i = 0; i < nx; ++i) { pr[2 * i] = ny * nnx + i; pr[2 * i + 1] = ny * nnx + i + 1; }
  if (phb_mesh_add_patch_by_nodes(m, "y+", nx, pr.data()) < 0) return PHB_ERR_ARG;
  return PHB_OK;
}

// device numbering + SELL pattern + SoA upload
int upload(phb_mesh *m) {
  cudaStream_t st = m->ctx->stream;
  const bool hostOnly = m->ctx->device < 0;  // numbering is still built; nothing is uploaded
  const int N = m->nCells, F = m->nFaces, nL = m->nLocal, P = m->nProcs;
  // device order: owned cells by IndexMap local row, then ghosts grouped by owner
  m->cell2dev.assign(N, -1);
  m->dev2cell.clear();
  m->dev2cell.resize(nL);
  for (int c = 0; c < N; ++c)
    if (m->owner[c] == m->rank) { m->cell2dev[c] = m->localRow[c]; m->dev2cell[m->localRow[c]] = c; }
  m->hRecvCnt.assign(P, 0); m->hRecvOff.assign(P, 0);
  m->hSendCnt.assign(P, 0); m->hSendOff.assign(P, 0);
  for (int q = 0; q < P; ++q) {
    m->hRecvOff[q] = (int)m->dev2cell.size();
    m->hRecvCnt[q] = m->bufPtr[q + 1] - m->bufPtr[q];
    for (int j = m->bufPtr[q]; j < m->bufPtr[q + 1]; ++j) {
      m->cell2dev[m->bufCell[j]] = (int)m->dev2cell.size();
      m->dev2cell.push_back(m->bufCell[j]);
    }
  }
  m->nDev = (int)m->dev2cell.size();
  m->hSendDev.clear();
  for (int q = 0; q < P; ++q) {
    m->hSendOff[q] = (int)m->hSendDev.size();
    m->hSendCnt[q] = m->sendPtr[q + 1] - m->sendPtr[q];
    for (int j = m->sendPtr[q]; j < m->sendPtr[q + 1]; ++j) m->hSendDev.push_back(m->cell2dev[m->sendCell[j]]);
  }
  // SELL pattern over owned rows
  SellPattern &S = m->sell;
  S.nRows = nL; S.nCols = m->nDev;
  S.nSlices = (nL + phb::kSliceRows - 1) / phb::kSliceRows;
  S.hRowLen.assign(nL, 0);
  S.hSliceOff.assign(S.nSlices + 1, 0);
  S.nnz = 0;
  for (int d = 0; d < nL; ++d) {
    const int c = m->dev2cell[d];
    S.hRowLen[d] = 1 + m->ilPtr[c + 1] - m->ilPtr[c];
    S.nnz += S.hRowLen[d];
  }
  for (int s = 0; s < S.nSlices; ++s) {
    int w = 0;
    for (int d = s * 32; d < std::min(nL, s * 32 + 32); ++d) w = std::max(w, S.hRowLen[d]);
    S.hSliceOff[s + 1] = S.hSliceOff[s] + w * 32;
  }
  S.nSlots = S.hSliceOff[S.nSlices];
  S.hCol.assign(S.nSlots, 0);
  std::vector<int> linkFace(S.nSlots, -1);
  for (int s = 0; s < S.nSlices; ++s) {
    const int w = (S.hSliceOff[s + 1] - S.hSliceOff[s]) / 32;
    for (int lane = 0; lane < 32; ++lane) {
      const int d = s * 32 + lane;
      for (int k = 0; k < w; ++k) {
        const size_t slot = (size_t)S.hSliceOff[s] + (size_t)k * 32 + lane;
        if (d >= nL) { S.hCol[slot] = nL ? nL - 1 : 0; continue; }
        const int c = m->dev2cell[d];
        if (k == 0 || k >= S.hRowLen[d]) { S.hCol[slot] = d; continue; }
        const int j = m->ilPtr[c] + k - 1;
        S.hCol[slot] = m->cell2dev[m->ilCell[j]];
        linkFace[slot] = m->ilFace[j] * 2 + (m->fR[m->ilFace[j]] == c ? 1 : 0);
      }
    }
  }
  if (hostOnly) return PHB_OK;
  PHB_CHECK(S.sliceOff.upload(S.hSliceOff, st));
  PHB_CHECK(S.rowLen.upload(S.hRowLen, st));
  PHB_CHECK(S.col.upload(S.hCol, st));
  PHB_CHECK(m->dLinkFace.upload(linkFace, st));
  // per-cell / per-face SoA
  std::vector<double> volDev(m->nDev);
  for (int d = 0; d < m->nDev; ++d) volDev[d] = m->vol[m->dev2cell[d]];
  PHB_CHECK(m->dVol.upload(volDev, st));
  PHB_CHECK(m->dFSx.upload(m->fSx, st)); PHB_CHECK(m->dFSy.upload(m->fSy, st));
  PHB_CHECK(m->dFG.upload(m->fG, st)); PHB_CHECK(m->dFW.upload(m->fW, st));
  PHB_CHECK(m->dFQx.upload(m->fQx, st)); PHB_CHECK(m->dFQy.upload(m->fQy, st));
  std::vector<int> fl(F), fr(F);
  for (int f = 0; f < F; ++f) { fl[f] = m->cell2dev[m->fL[f]]; fr[f] = m->fR[f] < 0 ? -1 : m->cell2dev[m->fR[f]]; }
  PHB_CHECK(m->dFL.upload(fl, st)); PHB_CHECK(m->dFR.upload(fr, st));
  // boundary cells (owned) and boundary / interior face lists
  std::vector<int> bcCell, bcPtr(1, 0), bcFace, bfFace, bfCell, bfPatch, ifFace;
  for (int d = 0; d < nL; ++d) {
    const int c = m->dev2cell[d];
    if (m->blPtr[c + 1] == m->blPtr[c]) continue;
    bcCell.push_back(d);
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) bcFace.push_back(m->blFace[j]);
    bcPtr.push_back((int)bcFace.size());
  }
  for (int f = 0; f < F; ++f) {
    if (m->fR[f] < 0) { bfFace.push_back(f); bfCell.push_back(m->cell2dev[m->fL[f]]); bfPatch.push_back(m->fPatch[f]); }
    else ifFace.push_back(f);
  }
  m->nBCells = (int)bcCell.size(); m->nBFaces = (int)bfFace.size(); m->nIFaces = (int)ifFace.size();
  PHB_CHECK(m->dBcCell.upload(bcCell, st)); PHB_CHECK(m->dBcPtr.upload(bcPtr, st));
  PHB_CHECK(m->dBcFace.upload(bcFace, st));
  PHB_CHECK(m->dBfFace.upload(bfFace, st)); PHB_CHECK(m->dBfCell.upload(bfCell, st));
  PHB_CHECK(m->dBfPatch.upload(bfPatch, st));
  PHB_CHECK(m->dIfFace.upload(ifFace, st));
  PHB_CHECK(m->dSendDev.upload(m->hSendDev, st));
  PHB_CHECK(m->dSendBuf.alloc(std::max<size_t>(1, 2 * m->hSendDev.size())));
  PHB_CHECK(m->dCell2Dev.upload(m->cell2dev, st));
  m->identityCells = m->nDev == m->nCells;
  for (int i = 0; i < m->nCells && m->identityCells; ++i) m->identityCells = m->cell2dev[i] == i;
  PHB_CUDA(cudaStreamSynchronize(st));
  return PHB_OK;
}


// ownership, IndexMap (1 index), buffer and send groups of a local mesh whose
// cells are the global cells `keep` (ascending).  UG/FiniteVolumeGrid2D.cpp:379-387,
// 460-511; UE/IndexMap.cpp:15-40.  The send group for peer q is what q keeps of
// ours, in q's local (= ascending global id) order, so no handshake is needed.
template <class PartOf, class Touches, class GRowOf>
void finish_local(phb_mesh *m, int rank, int P, const std::vector<int> &keep, PartOf partOf, Touches touches,
                  const std::vector<int> &offs, GRowOf gRowOf) {
  const int n = (int)keep.size();
  m->rank = rank; m->nProcs = P;
  m->nLocal = offs[rank + 1] - offs[rank]; m->rowOffset = offs[rank];
  m->owner.resize(n); m->globalId.resize(n); m->localRow.assign(n, -1); m->globalRow.resize(n);
  m->bufPtr.assign(P + 1, 0);
  for (int i = 0; i < n; ++i) {
    const int c = keep[i], q = partOf(c);
    m->owner[i] = q; m->globalId[i] = c; m->globalRow[i] = gRowOf(c);
    if (q == rank) m->localRow[i] = m->globalRow[i] - offs[rank];
    else m->bufPtr[q + 1]++;
  }
  std::partial_sum(m->bufPtr.begin(), m->bufPtr.end(), m->bufPtr.begin());
  m->bufCell.resize(m->bufPtr[P]);
  std::vector<int> fill(P, 0);
  for (int i = 0; i < n; ++i)
    if (m->owner[i] != rank) m->bufCell[m->bufPtr[m->owner[i]] + fill[m->owner[i]]++] = i;
  m->sendPtr.assign(P + 1, 0);
  m->sendCell.clear();
  for (int q = 0; q < P; ++q) {
    if (q != rank)
      for (int i = 0; i < n; ++i)
        if (m->owner[i] == rank && touches(keep[i], q)) m->sendCell.push_back(i);
    m->sendPtr[q + 1] = (int)m->sendCell.size();
  }
}

}  // namespace

extern "C" {

int phb_mesh_create(phb_ctx *ctx, int nNodes, const double *xy, int nCells, const int *cptr,
                    const int *cind, phb_mesh **out) {
  PHB_TRY_BEGIN
  return mesh_from_arrays(ctx, nNodes, xy, nCells, cptr, cind, out);
  PHB_TRY_END
}

int phb_mesh_create_rectilinear(phb_ctx *ctx, int nx, int ny, double w, double h, phb_mesh **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(nx > 0 && ny > 0 && w > 0 && h > 0, "phb_mesh_create_rectilinear: bad size");
  PHB_REQUIRE((long long)nx * ny < (1LL << 28), "phb_mesh_create_rectilinear: too many cells for int32 slots");
  std::vector<double> xy;
  rect_nodes(nx, ny, w, h, xy);
  const int nnx = nx + 1;
  std::vector<int> cptr((size_t)nx * ny + 1), cind(4 * (size_t)nx * ny);
  size_t c = 0;
  cptr[0] = 0;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i, ++c) {
      cptr[c + 1] = (int)(4 * (c + 1));
      cind[4 * c] = j * nnx + i;           // bl, br, tr, tl
      cind[4 * c + 1] = j * nnx + i + 1;
      cind[4 * c + 2] = (j + 1) * nnx + i + 1;
      cind[4 * c + 3] = (j + 1) * nnx + i;
    }
  PHB_CHECK(mesh_from_arrays(ctx, (nx + 1) * (ny + 1), xy.data(), nx * ny, cptr.data(), cind.data(), out));
  return rect_patches(*out, nx, ny);
  PHB_TRY_END
}

int phb_mesh_create_triangulated(phb_ctx *ctx, int nx, int ny, double w, double h, phb_mesh **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(nx > 0 && ny > 0 && w > 0 && h > 0, "phb_mesh_create_triangulated: bad size");
  std::vector<double> xy;
  rect_nodes(nx, ny, w, h, xy);
  const int nnx = nx + 1;
  const size_t nT = 2 * (size_t)nx * ny;
  std::vector<int> cptr(nT + 1), cind(3 * nT);
  size_t c = 0;
  cptr[0] = 0;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      const int bl = j * nnx + i, br = bl + 1, tl = (j + 1) * nnx + i, tr = tl + 1;
      int t[6];
      if ((i + j) % 2 == 0) { t[0] = bl; t[1] = br; t[2] = tr; t[3] = bl; t[4] = tr; t[5] = tl; }
      else { t[0] = bl; t[1] = br; t[2] = tl; t[3] = br; t[4] = tr; t[5] = tl; }
      for (int k = 0; k < 2; ++k, ++c) {
        cptr[c + 1] = (int)(3 * (c + 1));
        for (int q = 0; q < 3; ++q) cind[3 * c + q] = t[3 * k + q];
      }
    }
  PHB_CHECK(mesh_from_arrays(ctx, (nx + 1) * (ny + 1), xy.data(), (int)nT, cptr.data(), cind.data(), out));
  return rect_patches(*out, nx, ny);
  PHB_TRY_END
}

int phb_mesh_add_patch_by_nodes(phb_mesh *m, const char *name, int nPairs, const int *pairs) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(m && name && (pairs || nPairs == 0), "phb_mesh_add_patch_by_nodes: NULL argument");
  PHB_REQUIRE(!m->finalized, "phb_mesh_add_patch_by_nodes: mesh already finalized");
  int id = phb_mesh_patch_id(m, name);
  if (id < 0) {
    id = (int)m->patchNames.size();
    m->patchNames.push_back(name);
  }
  for (int i = 0; i < nPairs; ++i) {
    const int f = dir_find(m, pairs[2 * i], pairs[2 * i + 1]);
    PHB_REQUIRE(f >= 0, "no face found between n1 = %d, n2 = %d", pairs[2 * i], pairs[2 * i + 1]);
    m->fPatch[f] = id;
  }
  return id;
  PHB_TRY_END
}

int phb_mesh_patch_id(const phb_mesh *m, const char *name) {
  if (!m || !name) return -1;
  for (size_t i = 0; i < m->patchNames.size(); ++i)
    if (m->patchNames[i] == name) return (int)i;
  return -1;
}

int phb_mesh_finalize(phb_mesh *m) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(m, "phb_mesh_finalize: mesh is NULL");
  if (m->finalized) return PHB_OK;
  PHB_CHECK(build_geometry(m));
  build_links(m);
  // canonical CSR [P, nb...] over owned rows (reference: compact layout of any
  // operator sum, M/CrsEquation.cpp:185-275) + face -> slot map
  const int N = m->nCells, nL = m->nLocal;
  std::vector<int> rowCell(nL);
  for (int c = 0; c < N; ++c)
    if (m->owner[c] == m->rank) rowCell[m->localRow[c]] = c;
  m->rowPtr.assign(nL + 1, 0);
  for (int r = 0; r < nL; ++r) {
    const int c = rowCell[r];
    m->rowPtr[r + 1] = m->rowPtr[r] + 1 + m->ilPtr[c + 1] - m->ilPtr[c];
  }
  m->colInd.resize(m->rowPtr[nL]);
  m->slotL.assign(m->nFaces, -1); m->slotR.assign(m->nFaces, -1);
  m->slotDiag.resize(nL);
  for (int r = 0; r < nL; ++r) {
    const int c = rowCell[r];
    int k = m->rowPtr[r];
    m->slotDiag[r] = k;
    m->colInd[k++] = m->globalRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j, ++k) {
      m->colInd[k] = m->globalRow[m->ilCell[j]];
      const int f = m->ilFace[j];
      if (m->fL[f] == c) m->slotL[f] = k; else m->slotR[f] = k;
    }
  }
  m->nBFaces = 0;   // also on host-only contexts, where upload() has nothing to do
  for (int f = 0; f < m->nFaces; ++f) m->nBFaces += m->fR[f] < 0;
  PHB_CHECK(upload(m));
  m->finalized = true;
  return PHB_OK;
  PHB_TRY_END
}

int phb_mesh_set_peer_layout(phb_mesh *m, const int *peerRecvOff, const int *peerLd) {
  PHB_REQUIRE(m && peerRecvOff && peerLd, "phb_mesh_set_peer_layout: NULL argument");
  m->peerRecvOff.assign(peerRecvOff, peerRecvOff + m->nProcs);
  m->peerLd.assign(peerLd, peerLd + m->nProcs);
  return PHB_OK;
}

int phb_mesh_destroy(phb_mesh *m) {
  delete m;
  return PHB_OK;
}

int phb_mesh_sizes(const phb_mesh *m, long long out[11]) {
  PHB_REQUIRE(m && out, "phb_mesh_sizes: NULL argument");
  long long nb = m->nBFaces;   // counted once by finalize
  if (!m->finalized) { nb = 0; for (int f = 0; f < m->nFaces; ++f) nb += m->fR[f] < 0; }
  out[0] = m->nNodes; out[1] = m->nCells; out[2] = m->nFaces; out[3] = (long long)m->patchNames.size();
  out[4] = m->rank; out[5] = m->nProcs; out[6] = m->nLocal; out[7] = m->rowOffset;
  out[8] = m->nFaces - nb; out[9] = nb; out[10] = m->finalized ? (long long)m->rowPtr.back() : -1;
  return PHB_OK;
}

#define GET(nm, vec)                                                              \
  if (!strcmp(name, nm)) {                                                        \
    const long long n = (long long)(vec).size();                                  \
    if (out) std::copy((vec).begin(), (vec).begin() + std::min(n, cap), out);     \
    return n;                                                                     \
  }

long long phb_mesh_get_i32(const phb_mesh *m, const char *name, int *out, long long cap) {
  if (!m || !name) return PHB_ERR_ARG;
  GET("cptr", m->cptr) GET("cind", m->cind) GET("faceN1", m->fN1) GET("faceN2", m->fN2)
  GET("faceL", m->fL) GET("faceR", m->fR) GET("facePatch", m->fPatch)
  GET("ilPtr", m->ilPtr) GET("ilFace", m->ilFace) GET("ilCell", m->ilCell)
  GET("blPtr", m->blPtr) GET("blFace", m->blFace) GET("dlPtr", m->dlPtr) GET("dlCell", m->dlCell)
  GET("rowPtr", m->rowPtr) GET("colInd", m->colInd) GET("slotL", m->slotL) GET("slotR", m->slotR)
  GET("slotDiag", m->slotDiag) GET("owner", m->owner) GET("globalId", m->globalId)
  GET("localRow", m->localRow) GET("globalRow", m->globalRow) GET("bufPtr", m->bufPtr)
  GET("bufCell", m->bufCell) GET("sendPtr", m->sendPtr) GET("sendCell", m->sendCell)
  GET("cell2dev", m->cell2dev) GET("recvOff", m->hRecvOff) GET("recvCnt", m->hRecvCnt) GET("sellSliceOff", m->sell.hSliceOff) GET("sellCol", m->sell.hCol)
  GET("sellRowLen", m->sell.hRowLen)
  phb::set_error("phb_mesh_get_i32: unknown array \"%s\"", name);
  return PHB_ERR_ARG;
}
long long phb_mesh_get_f64(const phb_mesh *m, const char *name, double *out, long long cap) {
  if (!m || !name) return PHB_ERR_ARG;
  GET("nodeX", m->nodeX) GET("nodeY", m->nodeY) GET("vol", m->vol) GET("cellCx", m->cCx)
  GET("cellCy", m->cCy) GET("faceCx", m->fCx) GET("faceCy", m->fCy) GET("faceSx", m->fSx)
  GET("faceSy", m->fSy) GET("faceG", m->fG) GET("faceW", m->fW) GET("faceQx", m->fQx)
  GET("faceQy", m->fQy)
  phb::set_error("phb_mesh_get_f64: unknown array \"%s\"", name);
  return PHB_ERR_ARG;
}
#undef GET

// ------------------------------------------------------------ partition (I5)
int phb_partition_rcb(const phb_mesh *g, int nParts, int *part) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(g && part && nParts >= 1, "phb_partition_rcb: bad argument");
  const int N = g->nCells;
  std::vector<int> ids(N);
  std::iota(ids.begin(), ids.end(), 0);
  struct Job { int b, e, p0, np; };
  std::vector<Job> stack{{0, N, 0, nParts}};
  while (!stack.empty()) {
    Job j = stack.back();
    stack.pop_back();
    if (j.np == 1) {
      for (int i = j.b; i < j.e; ++i) part[ids[i]] = j.p0;
      continue;
    }
    double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
    for (int i = j.b; i < j.e; ++i) {
      const int c = ids[i];
      x0 = std::min(x0, g->cCx[c]); x1 = std::max(x1, g->cCx[c]);
      y0 = std::min(y0, g->cCy[c]); y1 = std::max(y1, g->cCy[c]);
    }
    const bool alongX = (x1 - x0) >= (y1 - y0);
    const std::vector<double> &co = alongX ? g->cCx : g->cCy;
    const int np1 = j.np / 2;
    const int n1 = (int)(((long long)(j.e - j.b) * np1) / j.np);
    auto cmp = [&](int a, int b) { return co[a] < co[b] || (co[a] == co[b] && a < b); };
    std::nth_element(ids.begin() + j.b, ids.begin() + j.b + n1, ids.begin() + j.e, cmp);
    stack.push_back({j.b, j.b + n1, j.p0, np1});
    stack.push_back({j.b + n1, j.e, j.p0 + np1, j.np - np1});
  }
  return PHB_OK;
  PHB_TRY_END
}

int phb_mesh_create_local(phb_ctx *ctx, const phb_mesh *g, const int *part, phb_mesh **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(ctx && g && part && out, "phb_mesh_create_local: NULL argument");
  PHB_REQUIRE(g->finalized, "phb_mesh_create_local: global mesh must be finalized");
  const int rank = ctx->rank, P = ctx->nProcs, N = g->nCells;
  auto touches = [&](int c, int q) {
    if (part[c] == q) return true;
    for (int j = g->ilPtr[c]; j < g->ilPtr[c + 1]; ++j) if (part[g->ilCell[j]] == q) return true;
    for (int j = g->dlPtr[c]; j < g->dlPtr[c + 1]; ++j) if (part[g->dlCell[j]] == q) return true;
    return false;
  };
  std::vector<int> keep, localNode(g->nNodes, -1), cptr(1, 0), cind, g2l(N, -1);
  std::vector<double> xy;
  for (int c = 0; c < N; ++c) {
    if (!touches(c, rank)) continue;
    g2l[c] = (int)keep.size();
    keep.push_back(c);
    for (int k = g->cptr[c]; k < g->cptr[c + 1]; ++k) {
      const int n = g->cind[k];
      if (localNode[n] < 0) {
        localNode[n] = (int)(xy.size() / 2);
        xy.push_back(g->nodeX[n]); xy.push_back(g->nodeY[n]);
      }
      cind.push_back(localNode[n]);
    }
    cptr.push_back((int)cind.size());
  }
  PHB_REQUIRE(!keep.empty(), "phb_mesh_create_local: rank %d owns no cells", rank);
  phb_mesh *m = nullptr;
  PHB_CHECK(mesh_from_arrays(ctx, (int)(xy.size() / 2), xy.data(), (int)keep.size(), cptr.data(), cind.data(), &m));
  std::unique_ptr<phb_mesh> guard(m);
  for (size_t p = 0; p < g->patchNames.size(); ++p) {
    std::vector<int> pr;
    for (int f = 0; f < g->nFaces; ++f) {
      if (g->fPatch[f] != (int)p) continue;
      const int a = localNode[g->fN1[f]], b = localNode[g->fN2[f]];
      if (a < 0 || b < 0 || dir_find(m, a, b) < 0) continue;
      pr.push_back(a); pr.push_back(b);
    }
    if (!pr.empty() && phb_mesh_add_patch_by_nodes(m, g->patchNames[p].c_str(), (int)pr.size() / 2, pr.data()) < 0)
      return PHB_ERR_ARG;
  }
  std::vector<int> nLocalOf(P, 0);
  for (int c = 0; c < N; ++c) nLocalOf[part[c]]++;
  std::vector<int> offs(P + 1, 0);
  for (int q = 0; q < P; ++q) offs[q + 1] = offs[q] + nLocalOf[q];
  // global row of every global cell: owner offset + rank among the owner's cells (ascending id)
  std::vector<int> gRow(N), cnt(P, 0);
  for (int c = 0; c < N; ++c) gRow[c] = offs[part[c]] + cnt[part[c]]++;
  finish_local(m, rank, P, keep, [&](int c) { return part[c]; }, touches, offs, [&](int c) { return gRow[c]; });
  PHB_CHECK(phb_mesh_finalize(m));
  *out = guard.release();
  return PHB_OK;
  PHB_TRY_END
}

// Local mesh of a px x py block partition of an nx x ny rectilinear grid, built
// WITHOUT materialising the global mesh (weak-scaling runs): rank q = bj px + bi owns
// columns [bi nx/px, (bi+1) nx/px) x rows [bj ny/py, (bj+1) ny/py); the result is
// identical (numbering, patches, halo maps) to phb_mesh_create_rectilinear + that
// partition vector + phb_mesh_create_local.
int phb_mesh_create_rect_block(phb_ctx *ctx, int nx, int ny, double w, double h, int px, int py,
                               phb_mesh **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(ctx && out && nx > 0 && ny > 0 && w > 0 && h > 0, "phb_mesh_create_rect_block: bad argument");
  const int rank = ctx->rank, P = ctx->nProcs;
  PHB_REQUIRE(px >= 1 && py >= 1 && px * py == P, "phb_mesh_create_rect_block: px*py (%d*%d) != nProcs %d", px, py, P);
  PHB_REQUIRE(ny >= py && nx >= px, "phb_mesh_create_rect_block: fewer rows/columns than blocks");
  auto rowBlock = [&](int j) { return (int)(((long long)j * py) / ny); };
  auto colBlock = [&](int i) { return (int)(((long long)i * px) / nx); };
  auto firstRow = [&](int b) { return b >= py ? ny : (int)(((long long)b * ny + py - 1) / py); };
  auto firstCol = [&](int b) { return b >= px ? nx : (int)(((long long)b * nx + px - 1) / px); };
  const int bi = rank % px, bj = rank / px;
  const int i0 = firstCol(bi), i1 = firstCol(bi + 1), j0 = firstRow(bj), j1 = firstRow(bj + 1);
  // kept cells: the block plus its face/node neighbours (one ring)
  const int ilo = std::max(0, i0 - 1), ihi = std::min(nx, i1 + 1), jlo = std::max(0, j0 - 1), jhi = std::min(ny, j1 + 1);
  const double hx0 = w / nx, hy0 = h / ny;
  const int lnx = ihi - ilo + 1;  // local node columns
  std::vector<int> keep, cptr(1, 0), cind, localNode((size_t)(jhi - jlo + 1) * lnx, -1);
  std::vector<double> xy;
  keep.reserve((size_t)(jhi - jlo) * (ihi - ilo));
  auto nodeId = [&](int i, int j) -> int {
    int &ln = localNode[(size_t)(j - jlo) * lnx + (i - ilo)];
    if (ln < 0) {
      ln = (int)(xy.size() / 2);
      xy.push_back(i * hx0); xy.push_back(j * hy0);
    }
    return ln;
  };
  for (int j = jlo; j < jhi; ++j)
    for (int i = ilo; i < ihi; ++i) {
      keep.push_back(j * nx + i);
      cind.push_back(nodeId(i, j)); cind.push_back(nodeId(i + 1, j));
      cind.push_back(nodeId(i + 1, j + 1)); cind.push_back(nodeId(i, j + 1));
      cptr.push_back((int)cind.size());
    }
  phb_mesh *m = nullptr;
  PHB_CHECK(mesh_from_arrays(ctx, (int)(xy.size() / 2), xy.data(), (int)keep.size(), cptr.data(), cind.data(), &m));
  std::unique_ptr<phb_mesh> guard(m);
  auto ln = [&](int i, int j) { return localNode[(size_t)(j - jlo) * lnx + (i - ilo)]; };
  std::vector<int> pr;
  // same creation order as the global grid: x-, x+, y-, y+ (a patch absent locally is not created)
  pr.clear();
  if (ilo == 0) for (int j = jlo; j < jhi; ++j) { pr.push_back(ln(0, j)); pr.push_back(ln(0, j + 1)); }
  if (!pr.empty() && phb_mesh_add_patch_by_nodes(m, "x-", (int)pr.size() / 2, pr.data()) < 0) return PHB_ERR_ARG;
  pr.clear();
  if (ihi == nx) for (int j = jlo; j < jhi; ++j) { pr.push_back(ln(nx, j)); pr.push_back(ln(nx, j + 1)); }
  if (!pr.empty() && phb_mesh_add_patch_by_nodes(m, "x+", (int)pr.size() / 2, pr.data()) < 0) return PHB_ERR_ARG;
  pr.clear();
  if (jlo == 0) for (int i = ilo; i < ihi; ++i) { pr.push_back(ln(i, 0)); pr.push_back(ln(i + 1, 0)); }
  if (!pr.empty() && phb_mesh_add_patch_by_nodes(m, "y-", (int)pr.size() / 2, pr.data()) < 0) return PHB_ERR_ARG;
  pr.clear();
  if (jhi == ny) for (int i = ilo; i < ihi; ++i) { pr.push_back(ln(i, ny)); pr.push_back(ln(i + 1, ny)); }
  if (!pr.empty() && phb_mesh_add_patch_by_nodes(m, "y+", (int)pr.size() / 2, pr.data()) < 0) return PHB_ERR_ARG;
  // IndexMap offsets: owned rows contiguous per rank, in rank order
  std::vector<int> offs(P + 1, 0);
  for (int q = 0; q < P; ++q) {
    const int qi = q % px, qj = q / px;
    offs[q + 1] = offs[q] + (firstCol(qi + 1) - firstCol(qi)) * (firstRow(qj + 1) - firstRow(qj));
  }
  auto partOf = [&](int c) { return rowBlock(c / nx) * px + colBlock(c % nx); };
  auto touches = [&](int c, int q) {
    const int i = c % nx, j = c / nx;
    for (int jj = std::max(0, j - 1); jj <= std::min(ny - 1, j + 1); ++jj)
      for (int ii = std::max(0, i - 1); ii <= std::min(nx - 1, i + 1); ++ii)
        if (rowBlock(jj) * px + colBlock(ii) == q) return true;
    return false;
  };
  // global row = owner offset + rank of the cell among the owner's cells in ascending global id
  auto gRowOf = [&](int c) {
    const int i = c % nx, j = c / nx, qi = colBlock(i), qj = rowBlock(j);
    const int bw = firstCol(qi + 1) - firstCol(qi);
    return offs[qj * px + qi] + (j - firstRow(qj)) * bw + (i - firstCol(qi));
  };
  finish_local(m, rank, P, keep, partOf, touches, offs, gRowOf);
  PHB_CHECK(phb_mesh_finalize(m));
  *out = guard.release();
  return PHB_OK;
  PHB_TRY_END
}

int phb_mesh_create_rect_strip(phb_ctx *ctx, int nx, int ny, double w, double h, phb_mesh **out) {
  PHB_REQUIRE(ctx, "phb_mesh_create_rect_strip: ctx is NULL");
  return phb_mesh_create_rect_block(ctx, nx, ny, w, h, 1, ctx->nProcs, out);
}

}  // extern "C"
