/*
 * phase_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY; see phase_oracle.h).
 *
 * Flat-array restatement of the reference's hot path.  Every function cites the
 * reference file:line it follows (paths relative to /root/reference/src; UG =
 * 2D/Unstructured/FiniteVolumeGrid2D, UF = 2D/Unstructured/FiniteVolume/Field,
 * UD = .../FiniteVolume/Discretization, UE = .../FiniteVolume/Equation,
 * US = 2D/Unstructured/Solvers, M = Math).
 *
 * Parity: pinned for the CSR algebra by oracle/_ref (reference sources compiled
 * in place); "parity unpinned" for everything the reference's tests do not pin
 * (mesh, operators, time step: the code is the specification) and for the
 * solve arithmetic (Eigen/Trilinos are not vendored).
 */
#include "phase_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OR_MAX_PATCHES 64

struct OrMesh {
  int nNodes, nCells, nFaces;
  double *nodeX, *nodeY;
  int *cptr, *cind;
  /* faces, id = order of first appearance (UG/FiniteVolumeGrid2D.cpp:84-114) */
  int *fN1, *fN2, *fL, *fR, *fPatch;
  double *fCx, *fCy, *fNx, *fNy;
  /* cells */
  double *vol, *cCx, *cCy;
  /* interior links per cell, ascending face id (UG/...Grid2D.cpp:396-419) */
  int *ilPtr, *ilFace, *ilCell;
  double *ilRcx, *ilRcy, *ilSx, *ilSy, *ilRfx, *ilRfy;
  /* boundary links */
  int *blPtr, *blFace;
  double *blRfx, *blRfy, *blSx, *blSy;
  /* diagonal links (UG/...Grid2D.cpp:422-429) */
  int *dlPtr, *dlCell;
  /* patches */
  int nPatches;
  char patchName[OR_MAX_PATCHES][32];
  /* face directory (sorted node pair -> face id) */
  uint64_t *hKey;
  int *hVal;
  size_t hCap;
  /* parallel (I3, I5) */
  int rank, nProcs;
  int *owner, *globalId;      /* per local cell */
  int *localRow;              /* IndexMap::local(cell,0), -1 for ghosts */
  int *globalRow;             /* IndexMap::global(cell,0) */
  int nLocal;                 /* owned cells */
  int rowOffset;              /* ownershipRange.first for 1 index */
  int *bufPtr, *bufCell;      /* bufferCellGroups_[q]: CSR over procs */
  int *sendPtr, *sendCell;    /* sendCellGroups_[q] */
};

/* ------------------------------------------------------------------ utils */
static void *xcalloc(size_t n, size_t s) {
  void *p = calloc(n ? n : 1, s);
  if (!p) {
    fprintf(stderr, "phase_oracle: out of memory\n");
    abort();
  }
  return p;
}

static uint64_t pair_key(int a, int b) {
  uint32_t lo = (uint32_t)(a < b ? a : b), hi = (uint32_t)(a < b ? b : a);
  return ((uint64_t)lo << 32) | hi;
}
static size_t hash64(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (size_t)k;
}
static int dir_find(const OrMesh *m, int a, int b) {
  uint64_t key = pair_key(a, b);
  size_t i = hash64(key) & (m->hCap - 1);
  while (m->hVal[i] >= 0) {
    if (m->hKey[i] == key) return m->hVal[i];
    i = (i + 1) & (m->hCap - 1);
  }
  return -1;
}
static void dir_insert(OrMesh *m, int a, int b, int id) {
  uint64_t key = pair_key(a, b);
  size_t i = hash64(key) & (m->hCap - 1);
  while (m->hVal[i] >= 0) i = (i + 1) & (m->hCap - 1);
  m->hKey[i] = key;
  m->hVal[i] = id;
}

/* Face::outwardNorm(pt): UG/Face/Face.cpp:48-50 */
static void outward_norm(const OrMesh *m, int f, double px, double py,
                         double *sx, double *sy) {
  double d = (m->fCx[f] - px) * m->fNx[f] + (m->fCy[f] - py) * m->fNy[f];
  if (d > 0.) {
    *sx = m->fNx[f];
    *sy = m->fNy[f];
  } else {
    *sx = -m->fNx[f];
    *sy = -m->fNy[f];
  }
}

/* ------------------------------------------------------------------- mesh */
/* FiniteVolumeGrid2D::init(nodes,cptr,cind) + createCell + init():
 * UG/FiniteVolumeGrid2D.cpp:20-35, 84-114, 396-450; Face ctor UG/Face/Face.cpp:9-18;
 * Cell ctor UG/Cell/Cell.cpp:8-29 (polygon area/centroid, G/Polygon.cpp:218-232);
 * links UG/Link/{InteriorLink,BoundaryLink,CellLink}.cpp. */
OrMesh *or_mesh_create(int nNodes, const double *xy, int nCells,
                       const int *cptr, const int *cind) {
  OrMesh *m = (OrMesh *)xcalloc(1, sizeof(OrMesh));
  m->nNodes = nNodes;
  m->nCells = nCells;
  m->nodeX = (double *)xcalloc(nNodes, sizeof(double));
  m->nodeY = (double *)xcalloc(nNodes, sizeof(double));
  for (int i = 0; i < nNodes; ++i) {
    m->nodeX[i] = xy[2 * i];
    m->nodeY[i] = xy[2 * i + 1];
  }
  m->cptr = (int *)xcalloc(nCells + 1, sizeof(int));
  memcpy(m->cptr, cptr, (nCells + 1) * sizeof(int));
  int nInd = cptr[nCells];
  m->cind = (int *)xcalloc(nInd, sizeof(int));
  memcpy(m->cind, cind, nInd * sizeof(int));

  int maxFaces = nInd; /* upper bound */
  m->fN1 = (int *)xcalloc(maxFaces, sizeof(int));
  m->fN2 = (int *)xcalloc(maxFaces, sizeof(int));
  m->fL = (int *)xcalloc(maxFaces, sizeof(int));
  m->fR = (int *)xcalloc(maxFaces, sizeof(int));
  m->hCap = 16;
  while (m->hCap < (size_t)maxFaces * 2 + 16) m->hCap <<= 1;
  m->hKey = (uint64_t *)xcalloc(m->hCap, sizeof(uint64_t));
  m->hVal = (int *)xcalloc(m->hCap, sizeof(int));
  for (size_t i = 0; i < m->hCap; ++i) m->hVal[i] = -1;

  m->vol = (double *)xcalloc(nCells, sizeof(double));
  m->cCx = (double *)xcalloc(nCells, sizeof(double));
  m->cCy = (double *)xcalloc(nCells, sizeof(double));

  int nF = 0;
  for (int c = 0; c < nCells; ++c) {
    int b = cptr[c], e = cptr[c + 1], nv = e - b;
    /* polygon area + centroid (shoelace about the first vertex) */
    double x0 = m->nodeX[cind[b]], y0 = m->nodeY[cind[b]];
    double a2 = 0., sx = 0., sy = 0.;
    for (int k = 0; k < nv; ++k) {
      int n1 = cind[b + k], n2 = cind[b + (k + 1) % nv];
      double x1 = m->nodeX[n1] - x0, y1 = m->nodeY[n1] - y0;
      double x2 = m->nodeX[n2] - x0, y2 = m->nodeY[n2] - y0;
      double ai = x1 * y2 - x2 * y1;
      a2 += ai;
      sx += ai * (x1 + x2);
      sy += ai * (y1 + y2);
    }
    m->vol[c] = 0.5 * a2;
    m->cCx[c] = sx / (3. * a2) + x0;
    m->cCy[c] = sy / (3. * a2) + y0;
    if (m->vol[c] < 0.) { /* boost::geometry::correct re-orients */
      m->vol[c] = -m->vol[c];
    }
    /* edges -> faces (createCell, :93-111) */
    for (int k = 0; k < nv; ++k) {
      int n1 = cind[b + k], n2 = cind[b + (k + 1) % nv];
      int f = dir_find(m, n1, n2);
      if (f < 0) {
        f = nF++;
        m->fN1[f] = n1;
        m->fN2[f] = n2;
        m->fL[f] = c;
        m->fR[f] = -1;
        dir_insert(m, n1, n2, f);
      } else {
        if (m->fR[f] >= 0) {
          fprintf(stderr, "phase_oracle: face shared by >2 cells\n");
          abort();
        }
        m->fR[f] = c;
      }
    }
  }
  m->nFaces = nF;
  m->fPatch = (int *)xcalloc(nF, sizeof(int));
  m->fCx = (double *)xcalloc(nF, sizeof(double));
  m->fCy = (double *)xcalloc(nF, sizeof(double));
  m->fNx = (double *)xcalloc(nF, sizeof(double));
  m->fNy = (double *)xcalloc(nF, sizeof(double));
  for (int f = 0; f < nF; ++f) {
    m->fPatch[f] = -1;
    double lx = m->nodeX[m->fN1[f]], ly = m->nodeY[m->fN1[f]];
    double rx = m->nodeX[m->fN2[f]], ry = m->nodeY[m->fN2[f]];
    m->fCx[f] = 0.5 * (lx + rx);
    m->fCy[f] = 0.5 * (ly + ry);
    double tx = rx - lx, ty = ry - ly;
    m->fNx[f] = ty; /* Vector2D::normalVec = (y,-x) */
    m->fNy[f] = -tx;
  }

  /* links in ascending face id */
  m->ilPtr = (int *)xcalloc(nCells + 1, sizeof(int));
  m->blPtr = (int *)xcalloc(nCells + 1, sizeof(int));
  for (int f = 0; f < nF; ++f) {
    if (m->fR[f] < 0)
      m->blPtr[m->fL[f] + 1]++;
    else {
      m->ilPtr[m->fL[f] + 1]++;
      m->ilPtr[m->fR[f] + 1]++;
    }
  }
  for (int c = 0; c < nCells; ++c) {
    m->ilPtr[c + 1] += m->ilPtr[c];
    m->blPtr[c + 1] += m->blPtr[c];
  }
  int nIl = m->ilPtr[nCells], nBl = m->blPtr[nCells];
  m->ilFace = (int *)xcalloc(nIl, sizeof(int));
  m->ilCell = (int *)xcalloc(nIl, sizeof(int));
  m->ilRcx = (double *)xcalloc(nIl, sizeof(double));
  m->ilRcy = (double *)xcalloc(nIl, sizeof(double));
  m->ilSx = (double *)xcalloc(nIl, sizeof(double));
  m->ilSy = (double *)xcalloc(nIl, sizeof(double));
  m->ilRfx = (double *)xcalloc(nIl, sizeof(double));
  m->ilRfy = (double *)xcalloc(nIl, sizeof(double));
  m->blFace = (int *)xcalloc(nBl, sizeof(int));
  m->blRfx = (double *)xcalloc(nBl, sizeof(double));
  m->blRfy = (double *)xcalloc(nBl, sizeof(double));
  m->blSx = (double *)xcalloc(nBl, sizeof(double));
  m->blSy = (double *)xcalloc(nBl, sizeof(double));
  int *ilFill = (int *)xcalloc(nCells, sizeof(int));
  int *blFill = (int *)xcalloc(nCells, sizeof(int));
  for (int f = 0; f < nF; ++f) {
    if (m->fR[f] < 0) {
      int c = m->fL[f];
      int j = m->blPtr[c] + blFill[c]++;
      m->blFace[j] = f;
      m->blRfx[j] = m->fCx[f] - m->cCx[c];
      m->blRfy[j] = m->fCy[f] - m->cCy[c];
      outward_norm(m, f, m->cCx[c], m->cCy[c], &m->blSx[j], &m->blSy[j]);
    } else {
      int cc[2] = {m->fL[f], m->fR[f]};
      for (int s = 0; s < 2; ++s) {
        int self = cc[s], nb = cc[1 - s];
        int j = m->ilPtr[self] + ilFill[self]++;
        m->ilFace[j] = f;
        m->ilCell[j] = nb;
        m->ilRcx[j] = m->cCx[nb] - m->cCx[self];
        m->ilRcy[j] = m->cCy[nb] - m->cCy[self];
        outward_norm(m, f, m->cCx[self], m->cCy[self], &m->ilSx[j],
                     &m->ilSy[j]);
        m->ilRfx[j] = m->fCx[f] - m->cCx[self];
        m->ilRfy[j] = m->fCy[f] - m->cCy[self];
      }
    }
  }
  free(ilFill);
  free(blFill);

  /* diagonal links: node -> cells in ascending cell id */
  int *npPtr = (int *)xcalloc(nNodes + 1, sizeof(int));
  for (int i = 0; i < nInd; ++i) npPtr[cind[i] + 1]++;
  for (int i = 0; i < nNodes; ++i) npPtr[i + 1] += npPtr[i];
  int *npCell = (int *)xcalloc(nInd, sizeof(int));
  int *npFill = (int *)xcalloc(nNodes, sizeof(int));
  for (int c = 0; c < nCells; ++c)
    for (int k = cptr[c]; k < cptr[c + 1]; ++k) {
      int n = cind[k];
      npCell[npPtr[n] + npFill[n]++] = c;
    }
  free(npFill);
  m->dlPtr = (int *)xcalloc(nCells + 1, sizeof(int));
  for (int pass = 0; pass < 2; ++pass) {
    int cnt = 0;
    for (int c = 0; c < nCells; ++c) {
      for (int k = cptr[c]; k < cptr[c + 1]; ++k) {
        int n = cind[k];
        for (int q = npPtr[n]; q < npPtr[n + 1]; ++q) {
          int kc = npCell[q];
          if (kc == c) continue;
          int share = 0;
          for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j)
            if (m->ilCell[j] == kc) share = 1;
          if (share) continue;
          if (pass) m->dlCell[cnt] = kc;
          cnt++;
        }
      }
      if (!pass) m->dlPtr[c + 1] = cnt;
    }
    if (!pass) m->dlCell = (int *)xcalloc(cnt, sizeof(int));
  }
  free(npPtr);
  free(npCell);

  /* single-process defaults (init(): :441-445; IndexMap UE/IndexMap.cpp:15-40) */
  m->rank = 0;
  m->nProcs = 1;
  m->owner = (int *)xcalloc(nCells, sizeof(int));
  m->globalId = (int *)xcalloc(nCells, sizeof(int));
  m->localRow = (int *)xcalloc(nCells, sizeof(int));
  m->globalRow = (int *)xcalloc(nCells, sizeof(int));
  for (int c = 0; c < nCells; ++c) {
    m->globalId[c] = c;
    m->localRow[c] = c;
    m->globalRow[c] = c;
  }
  m->nLocal = nCells;
  m->rowOffset = 0;
  m->bufPtr = (int *)xcalloc(2, sizeof(int));
  m->sendPtr = (int *)xcalloc(2, sizeof(int));
  m->bufCell = (int *)xcalloc(1, sizeof(int));
  m->sendCell = (int *)xcalloc(1, sizeof(int));
  return m;
}

void or_mesh_destroy(OrMesh *m) {
  if (!m) return;
  void *ps[] = {m->nodeX, m->nodeY, m->cptr,  m->cind,  m->fN1,   m->fN2,
                m->fL,    m->fR,    m->fPatch, m->fCx,  m->fCy,   m->fNx,
                m->fNy,   m->vol,   m->cCx,   m->cCy,   m->ilPtr, m->ilFace,
                m->ilCell, m->ilRcx, m->ilRcy, m->ilSx, m->ilSy,  m->ilRfx,
                m->ilRfy, m->blPtr, m->blFace, m->blRfx, m->blRfy, m->blSx,
                m->blSy,  m->dlPtr, m->dlCell, m->hKey, m->hVal,  m->owner,
                m->globalId, m->localRow, m->globalRow, m->bufPtr, m->bufCell,
                m->sendPtr, m->sendCell};
  for (size_t i = 0; i < sizeof(ps) / sizeof(ps[0]); ++i) free(ps[i]);
  free(m);
}

int or_mesh_patch_id(const OrMesh *m, const char *name) {
  for (int i = 0; i < m->nPatches; ++i)
    if (!strcmp(m->patchName[i], name)) return i;
  return -1;
}

/* FiniteVolumeGrid2D::createPatchByNodes: UG/FiniteVolumeGrid2D.cpp:182-198 */
int or_mesh_add_patch_by_nodes(OrMesh *m, const char *name, int nPairs,
                               const int *nodePairs) {
  int id = or_mesh_patch_id(m, name);
  if (id < 0) {
    if (m->nPatches >= OR_MAX_PATCHES) return -1;
    id = m->nPatches++;
    strncpy(m->patchName[id], name, 31);
  }
  for (int i = 0; i < nPairs; ++i) {
    int f = dir_find(m, nodePairs[2 * i], nodePairs[2 * i + 1]);
    if (f < 0) return -2; /* findFace throws */
    m->fPatch[f] = id;
  }
  return id;
}

/* StructuredRectilinearGrid::init + initPatches:
 * UG/StructuredRectilinearGrid.cpp:40-95, 175-194 */
static void rect_nodes(int nx, int ny, double width, double height,
                       double *xy) {
  double hx0 = width / nx, hy0 = height / ny;
  for (int j = 0; j <= ny; ++j)
    for (int i = 0; i <= nx; ++i) {
      xy[2 * (j * (nx + 1) + i)] = i * hx0;
      xy[2 * (j * (nx + 1) + i) + 1] = j * hy0;
    }
}
static void rect_patches(OrMesh *m, int nx, int ny) {
  int nnx = nx + 1;
  int *pairs = (int *)xcalloc(2 * (nx > ny ? nx : ny), sizeof(int));
  for (int j = 0; j < ny; ++j) {
    pairs[2 * j] = j * nnx;
    pairs[2 * j + 1] = (j + 1) * nnx;
  }
  or_mesh_add_patch_by_nodes(m, "x-", ny, pairs);
  for (int j = 0; j < ny; ++j) {
    pairs[2 * j] = j * nnx + nx;
    pairs[2 * j + 1] = (j + 1) * nnx + nx;
  }
  or_mesh_add_patch_by_nodes(m, "x+", ny, pairs);
  for (int i = 0; i < nx; ++i) {
    pairs[2 * i] = i;
    pairs[2 * i + 1] = i + 1;
  }
  or_mesh_add_patch_by_nodes(m, "y-", nx, pairs);
  for (int i = 0; i < nx; ++i) {
    pairs[2 * i] = ny * nnx + i;
    pairs[2 * i + 1] = ny * nnx + i + 1;
  }
  or_mesh_add_patch_by_nodes(m, "y+", nx, pairs);
  free(pairs);
}

OrMesh *or_mesh_rectilinear(int nx, int ny, double width, double height) {
  int nnx = nx + 1;
  double *xy = (double *)xcalloc(2 * (size_t)(nx + 1) * (ny + 1), sizeof(double));
  rect_nodes(nx, ny, width, height, xy);
  int *cptr = (int *)xcalloc((size_t)nx * ny + 1, sizeof(int));
  int *cind = (int *)xcalloc(4 * (size_t)nx * ny, sizeof(int));
  int c = 0;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i, ++c) {
      cptr[c + 1] = cptr[c] + 4;
      cind[4 * c] = j * nnx + i;
      cind[4 * c + 1] = j * nnx + i + 1;
      cind[4 * c + 2] = (j + 1) * nnx + i + 1;
      cind[4 * c + 3] = (j + 1) * nnx + i;
    }
  OrMesh *m = or_mesh_create((nx + 1) * (ny + 1), xy, nx * ny, cptr, cind);
  rect_patches(m, nx, ny);
  free(xy);
  free(cptr);
  free(cind);
  return m;
}

/* synthetic unstructured variant (config 2): not a reference generator */
OrMesh *or_mesh_triangulated(int nx, int ny, double width, double height) {
  int nnx = nx + 1;
  double *xy = (double *)xcalloc(2 * (size_t)(nx + 1) * (ny + 1), sizeof(double));
  rect_nodes(nx, ny, width, height, xy);
  size_t nT = 2 * (size_t)nx * ny;
  int *cptr = (int *)xcalloc(nT + 1, sizeof(int));
  int *cind = (int *)xcalloc(3 * nT, sizeof(int));
  int c = 0;
  for (int j = 0; j < ny; ++j)
    for (int i = 0; i < nx; ++i) {
      int bl = j * nnx + i, br = bl + 1, tl = (j + 1) * nnx + i, tr = tl + 1;
      int t[6];
      if ((i + j) % 2 == 0) {
        t[0] = bl; t[1] = br; t[2] = tr;
        t[3] = bl; t[4] = tr; t[5] = tl;
      } else {
        t[0] = bl; t[1] = br; t[2] = tl;
        t[3] = br; t[4] = tr; t[5] = tl;
      }
      for (int k = 0; k < 2; ++k, ++c) {
        cptr[c + 1] = cptr[c] + 3;
        memcpy(cind + 3 * c, t + 3 * k, 3 * sizeof(int));
      }
    }
  OrMesh *m = or_mesh_create((nx + 1) * (ny + 1), xy, (int)nT, cptr, cind);
  rect_patches(m, nx, ny);
  free(xy);
  free(cptr);
  free(cind);
  return m;
}

#define ARR_I(nm, p, n)            \
  if (!strcmp(name, nm)) {         \
    *ptr = (p);                    \
    *isDouble = 0;                 \
    return (long)(n);              \
  }
#define ARR_D(nm, p, n)            \
  if (!strcmp(name, nm)) {         \
    *ptr = (p);                    \
    *isDouble = 1;                 \
    return (long)(n);              \
  }

long or_mesh_array(const OrMesh *m, const char *name, const void **ptr,
                   int *isDouble) {
  int N = m->nCells, F = m->nFaces;
  static int scal[8];
  if (!strcmp(name, "sizes")) {
    scal[0] = m->nNodes; scal[1] = N; scal[2] = F; scal[3] = m->nPatches;
    scal[4] = m->rank; scal[5] = m->nProcs; scal[6] = m->nLocal;
    scal[7] = m->rowOffset;
    *ptr = scal; *isDouble = 0;
    return 8;
  }
  ARR_D("nodeX", m->nodeX, m->nNodes) ARR_D("nodeY", m->nodeY, m->nNodes)
  ARR_I("cptr", m->cptr, N + 1) ARR_I("cind", m->cind, m->cptr[N])
  ARR_I("faceN1", m->fN1, F) ARR_I("faceN2", m->fN2, F)
  ARR_I("faceL", m->fL, F) ARR_I("faceR", m->fR, F)
  ARR_I("facePatch", m->fPatch, F)
  ARR_D("faceCx", m->fCx, F) ARR_D("faceCy", m->fCy, F)
  ARR_D("faceNx", m->fNx, F) ARR_D("faceNy", m->fNy, F)
  ARR_D("vol", m->vol, N) ARR_D("cellCx", m->cCx, N) ARR_D("cellCy", m->cCy, N)
  ARR_I("ilPtr", m->ilPtr, N + 1) ARR_I("ilFace", m->ilFace, m->ilPtr[N])
  ARR_I("ilCell", m->ilCell, m->ilPtr[N])
  ARR_D("ilRcx", m->ilRcx, m->ilPtr[N]) ARR_D("ilRcy", m->ilRcy, m->ilPtr[N])
  ARR_D("ilSx", m->ilSx, m->ilPtr[N]) ARR_D("ilSy", m->ilSy, m->ilPtr[N])
  ARR_I("blPtr", m->blPtr, N + 1) ARR_I("blFace", m->blFace, m->blPtr[N])
  ARR_D("blRfx", m->blRfx, m->blPtr[N]) ARR_D("blRfy", m->blRfy, m->blPtr[N])
  ARR_D("blSx", m->blSx, m->blPtr[N]) ARR_D("blSy", m->blSy, m->blPtr[N])
  ARR_I("dlPtr", m->dlPtr, N + 1) ARR_I("dlCell", m->dlCell, m->dlPtr[N])
  ARR_I("owner", m->owner, N) ARR_I("globalId", m->globalId, N)
  ARR_I("localRow", m->localRow, N) ARR_I("globalRow", m->globalRow, N)
  ARR_I("bufPtr", m->bufPtr, m->nProcs + 1)
  ARR_I("bufCell", m->bufCell, m->bufPtr[m->nProcs])
  ARR_I("sendPtr", m->sendPtr, m->nProcs + 1)
  ARR_I("sendCell", m->sendCell, m->sendPtr[m->nProcs])
  return -1;
}

/* ------------------------------------------------- partition + halo (I5) */
/* FiniteVolumeGrid2D::partition, UG/FiniteVolumeGrid2D.cpp:309-387, with the
 * partition VECTOR as input (the METIS call itself is unpinned):
 * keep every cell owned by `rank` or with any cellLink (face or diagonal)
 * neighbour owned by `rank`; cells renumbered in ascending global id, nodes in
 * first-use order; patches rebuilt from node pairs; minBufferWidth = 0. */
OrMesh *or_mesh_partition_local(const OrMesh *g, const int *part, int rank,
                                int nProcs) {
  int N = g->nCells;
  int *localNode = (int *)xcalloc(g->nNodes, sizeof(int));
  for (int i = 0; i < g->nNodes; ++i) localNode[i] = -1;
  int *keepIds = (int *)xcalloc(N, sizeof(int));
  int nKeep = 0, nInd = 0;
  for (int c = 0; c < N; ++c) {
    int keep = part[c] == rank;
    for (int j = g->ilPtr[c]; !keep && j < g->ilPtr[c + 1]; ++j)
      keep = part[g->ilCell[j]] == rank;
    for (int j = g->dlPtr[c]; !keep && j < g->dlPtr[c + 1]; ++j)
      keep = part[g->dlCell[j]] == rank;
    if (keep) {
      keepIds[nKeep++] = c;
      nInd += g->cptr[c + 1] - g->cptr[c];
    }
  }
  double *xy = (double *)xcalloc(2 * (size_t)g->nNodes, sizeof(double));
  int *cptr = (int *)xcalloc(nKeep + 1, sizeof(int));
  int *cind = (int *)xcalloc(nInd, sizeof(int));
  int nn = 0;
  for (int i = 0; i < nKeep; ++i) {
    int c = keepIds[i];
    cptr[i + 1] = cptr[i];
    for (int k = g->cptr[c]; k < g->cptr[c + 1]; ++k) {
      int n = g->cind[k];
      if (localNode[n] < 0) {
        localNode[n] = nn;
        xy[2 * nn] = g->nodeX[n];
        xy[2 * nn + 1] = g->nodeY[n];
        nn++;
      }
      cind[cptr[i + 1]++] = localNode[n];
    }
  }
  OrMesh *m = or_mesh_create(nn, xy, nKeep, cptr, cind);
  /* patches: faces whose two nodes both exist locally (:356-369) */
  for (int p = 0; p < g->nPatches; ++p) {
    int cnt = 0;
    for (int f = 0; f < g->nFaces; ++f)
      if (g->fPatch[f] == p && localNode[g->fN1[f]] >= 0 &&
          localNode[g->fN2[f]] >= 0)
        cnt++;
    if (!cnt) continue;
    int *pairs = (int *)xcalloc(2 * cnt, sizeof(int));
    cnt = 0;
    for (int f = 0; f < g->nFaces; ++f)
      if (g->fPatch[f] == p && localNode[g->fN1[f]] >= 0 &&
          localNode[g->fN2[f]] >= 0) {
        /* a node pair present locally may still not be a local face when the
         * owning cell was not kept; the reference's findFace would throw.  Only
         * faces that exist locally are patched. */
        if (dir_find(m, localNode[g->fN1[f]], localNode[g->fN2[f]]) < 0)
          continue;
        pairs[2 * cnt] = localNode[g->fN1[f]];
        pairs[2 * cnt + 1] = localNode[g->fN2[f]];
        cnt++;
      }
    or_mesh_add_patch_by_nodes(m, g->patchName[p], cnt, pairs);
    free(pairs);
  }
  m->rank = rank;
  m->nProcs = nProcs;
  m->nLocal = 0;
  for (int i = 0; i < nKeep; ++i) {
    m->owner[i] = part[keepIds[i]];
    m->globalId[i] = keepIds[i];
    if (m->owner[i] == rank) m->nLocal++;
  }
  /* bufferCellGroups_[q]: ghosts owned by q in local id order (:472-478) */
  free(m->bufPtr);
  free(m->bufCell);
  m->bufPtr = (int *)xcalloc(nProcs + 1, sizeof(int));
  for (int i = 0; i < nKeep; ++i)
    if (m->owner[i] != rank) m->bufPtr[m->owner[i] + 1]++;
  for (int q = 0; q < nProcs; ++q) m->bufPtr[q + 1] += m->bufPtr[q];
  m->bufCell = (int *)xcalloc(m->bufPtr[nProcs], sizeof(int));
  int *fill = (int *)xcalloc(nProcs, sizeof(int));
  for (int i = 0; i < nKeep; ++i)
    if (m->owner[i] != rank) {
      int q = m->owner[i];
      m->bufCell[m->bufPtr[q] + fill[q]++] = i;
    }
  free(fill);
  free(localNode);
  free(keepIds);
  free(xy);
  free(cptr);
  free(cind);
  return m;
}

/* initCommBuffers send side (UG/FiniteVolumeGrid2D.cpp:480-510) + IndexMap
 * (UE/IndexMap.cpp:15-40) for nIndices = 1, for all ranks at once. */
int or_mesh_init_comm(OrMesh **L, int nProcs) {
  int *offset = (int *)xcalloc(nProcs + 1, sizeof(int));
  for (int r = 0; r < nProcs; ++r) offset[r + 1] = offset[r] + L[r]->nLocal;
  int maxG = 0;
  for (int r = 0; r < nProcs; ++r)
    for (int i = 0; i < L[r]->nCells; ++i)
      if (L[r]->globalId[i] + 1 > maxG) maxG = L[r]->globalId[i] + 1;
  int *g2l = (int *)xcalloc(maxG, sizeof(int));
  int *gRow = (int *)xcalloc(maxG, sizeof(int));
  for (int r = 0; r < nProcs; ++r) {
    OrMesh *m = L[r];
    m->rowOffset = offset[r];
    int li = 0;
    for (int i = 0; i < m->nCells; ++i) {
      if (m->owner[i] == r) {
        m->localRow[i] = li;
        m->globalRow[i] = offset[r] + li;
        gRow[m->globalId[i]] = m->globalRow[i];
        li++;
      } else {
        m->localRow[i] = -1;
        m->globalRow[i] = -1;
      }
    }
  }
  for (int r = 0; r < nProcs; ++r) {
    OrMesh *m = L[r];
    for (int i = 0; i < m->nCells; ++i) g2l[m->globalId[i]] = i;
    free(m->sendPtr);
    free(m->sendCell);
    m->sendPtr = (int *)xcalloc(nProcs + 1, sizeof(int));
    for (int q = 0; q < nProcs; ++q) {
      int cnt = 0;
      if (q != r) cnt = L[q]->bufPtr[r + 1] - L[q]->bufPtr[r];
      m->sendPtr[q + 1] = m->sendPtr[q] + cnt;
    }
    m->sendCell = (int *)xcalloc(m->sendPtr[nProcs], sizeof(int));
    for (int q = 0; q < nProcs; ++q) {
      if (q == r) continue;
      int k = m->sendPtr[q];
      for (int j = L[q]->bufPtr[r]; j < L[q]->bufPtr[r + 1]; ++j)
        m->sendCell[k++] = g2l[L[q]->globalId[L[q]->bufCell[j]]];
    }
    /* ghosts receive their owner's global row (grid.sendMessages(globalIndices_)) */
    for (int i = 0; i < m->nCells; ++i)
      if (m->owner[i] != r) m->globalRow[i] = gRow[m->globalId[i]];
  }
  free(offset);
  free(g2l);
  free(gRow);
  return 0;
}

/* ---------------------------------------------------------- CrsEquation */
struct OrCrs {
  int nRows;
  int *rowPtr;
  int *colInd;
  double *vals;
  int cap; /* allocated slots */
  double *rhs;
};

/* CrsEquation::CrsEquation(nRows,nnz): M/CrsEquation.cpp:9-14 */
OrCrs *or_crs_create(int nRows, int nnz) {
  OrCrs *e = (OrCrs *)xcalloc(1, sizeof(OrCrs));
  e->nRows = nRows;
  e->rowPtr = (int *)xcalloc(nRows + 1, sizeof(int));
  for (int i = 0; i <= nRows; ++i) e->rowPtr[i] = i * nnz;
  e->cap = nRows * nnz;
  e->colInd = (int *)xcalloc(e->cap, sizeof(int));
  e->vals = (double *)xcalloc(e->cap, sizeof(double));
  for (int i = 0; i < e->cap; ++i) e->colInd[i] = -1;
  e->rhs = (double *)xcalloc(nRows, sizeof(double));
  return e;
}
OrCrs *or_crs_clone(const OrCrs *s) {
  OrCrs *e = (OrCrs *)xcalloc(1, sizeof(OrCrs));
  e->nRows = s->nRows;
  int nnz = s->rowPtr[s->nRows];
  e->cap = nnz;
  e->rowPtr = (int *)xcalloc(s->nRows + 1, sizeof(int));
  memcpy(e->rowPtr, s->rowPtr, (s->nRows + 1) * sizeof(int));
  e->colInd = (int *)xcalloc(nnz, sizeof(int));
  memcpy(e->colInd, s->colInd, nnz * sizeof(int));
  e->vals = (double *)xcalloc(nnz, sizeof(double));
  memcpy(e->vals, s->vals, nnz * sizeof(double));
  e->rhs = (double *)xcalloc(s->nRows, sizeof(double));
  memcpy(e->rhs, s->rhs, s->nRows * sizeof(double));
  return e;
}
void or_crs_destroy(OrCrs *e) {
  if (!e) return;
  free(e->rowPtr);
  free(e->colInd);
  free(e->vals);
  free(e->rhs);
  free(e);
}
static void crs_insert_slot(OrCrs *e, int row, int col, double v) {
  /* vector::insert at rowPtr[row+1] + rowPtr shift: M/CrsEquation.cpp:125-130 */
  int nnz = e->rowPtr[e->nRows], pos = e->rowPtr[row + 1];
  if (nnz + 1 > e->cap) {
    e->cap = nnz + 1 + e->nRows;
    e->colInd = (int *)realloc(e->colInd, e->cap * sizeof(int));
    e->vals = (double *)realloc(e->vals, e->cap * sizeof(double));
  }
  memmove(e->colInd + pos + 1, e->colInd + pos, (nnz - pos) * sizeof(int));
  memmove(e->vals + pos + 1, e->vals + pos, (nnz - pos) * sizeof(double));
  e->colInd[pos] = col;
  e->vals[pos] = v;
  for (int r = row + 1; r <= e->nRows; ++r) e->rowPtr[r]++;
}
/* CrsEquation::addCoeff: M/CrsEquation.cpp:113-131 */
void or_crs_add_coeff(OrCrs *e, int row, int col, double v) {
  for (int j = e->rowPtr[row]; j < e->rowPtr[row + 1]; ++j) {
    if (e->colInd[j] == col) {
      e->vals[j] += v;
      return;
    } else if (e->colInd[j] == -1) {
      e->colInd[j] = col;
      e->vals[j] = v;
      return;
    }
  }
  crs_insert_slot(e, row, col, v);
}
/* CrsEquation::setCoeff: M/CrsEquation.cpp:133-151 */
void or_crs_set_coeff(OrCrs *e, int row, int col, double v) {
  for (int j = e->rowPtr[row]; j < e->rowPtr[row + 1]; ++j) {
    if (e->colInd[j] == col || e->colInd[j] == -1) {
      e->colInd[j] = col;
      e->vals[j] = v;
      return;
    }
  }
  crs_insert_slot(e, row, col, v);
}
void or_crs_add_rhs(OrCrs *e, int row, double v) { e->rhs[row] += v; }
/* scaleRow: M/CrsEquation.cpp:153-159 */
void or_crs_scale_row(OrCrs *e, int row, double v) {
  for (int j = e->rowPtr[row]; j < e->rowPtr[row + 1]; ++j) e->vals[j] *= v;
  e->rhs[row] *= v;
}
/* operator+= / operator-= compaction: M/CrsEquation.cpp:185-275 */
static void crs_merge(OrCrs *l, const OrCrs *r, double sign) {
  int n = l->nRows;
  int capNew = l->rowPtr[n] + r->rowPtr[n];
  int *tp = (int *)xcalloc(n + 1, sizeof(int));
  int *tc = (int *)xcalloc(capNew, sizeof(int));
  double *tv = (double *)xcalloc(capNew, sizeof(double));
  int k = 0;
  for (int row = 0; row < n; ++row) {
    int first = k;
    for (int j = l->rowPtr[row]; j < l->rowPtr[row + 1]; ++j) {
      if (l->vals[j] == 0. || l->colInd[j] < 0) continue;
      tc[k] = l->colInd[j];
      tv[k] = l->vals[j];
      k++;
    }
    for (int j = r->rowPtr[row]; j < r->rowPtr[row + 1]; ++j) {
      if (r->vals[j] == 0. || r->colInd[j] < 0) continue;
      int hit = -1;
      for (int q = first; q < k; ++q)
        if (tc[q] == r->colInd[j]) {
          hit = q;
          break;
        }
      if (hit >= 0) {
        if (sign > 0)
          tv[hit] += r->vals[j];
        else
          tv[hit] -= r->vals[j];
      } else {
        tc[k] = r->colInd[j];
        tv[k] = sign > 0 ? r->vals[j] : -r->vals[j];
        k++;
      }
    }
    tp[row + 1] = k;
  }
  free(l->rowPtr);
  free(l->colInd);
  free(l->vals);
  l->rowPtr = tp;
  l->colInd = tc;
  l->vals = tv;
  l->cap = capNew;
  for (int i = 0; i < n; ++i) {
    if (sign > 0)
      l->rhs[i] += r->rhs[i];
    else
      l->rhs[i] -= r->rhs[i];
  }
}
void or_crs_add_eq(OrCrs *l, const OrCrs *r) { crs_merge(l, r, 1.); }
void or_crs_sub_eq(OrCrs *l, const OrCrs *r) { crs_merge(l, r, -1.); }
void or_crs_sub_vec(OrCrs *l, const double *v) {
  for (int i = 0; i < l->nRows; ++i) l->rhs[i] -= v[i];
}
void or_crs_add_vec(OrCrs *l, const double *v) {
  for (int i = 0; i < l->nRows; ++i) l->rhs[i] += v[i];
}
void or_crs_scale(OrCrs *l, double s) {
  int nnz = l->rowPtr[l->nRows];
  for (int j = 0; j < nnz; ++j) l->vals[j] *= s;
  for (int i = 0; i < l->nRows; ++i) l->rhs[i] *= s;
}
int or_crs_rank(const OrCrs *e) { return e->nRows; }
int or_crs_nnz(const OrCrs *e) { return e->rowPtr[e->nRows]; }
void or_crs_export(const OrCrs *e, int *rowPtr, int *colInd, double *vals,
                   double *rhs) {
  int nnz = e->rowPtr[e->nRows];
  if (rowPtr) memcpy(rowPtr, e->rowPtr, (e->nRows + 1) * sizeof(int));
  if (colInd) memcpy(colInd, e->colInd, nnz * sizeof(int));
  if (vals) memcpy(vals, e->vals, nnz * sizeof(double));
  if (rhs) memcpy(rhs, e->rhs, e->nRows * sizeof(double));
}

/* ------------------------------------------------------- fractional step */
typedef struct {
  int type;
  double vx, vy;
} OrBc;

struct OrFracStep {
  OrMesh *m;
  double rho, mu;
  /* fields: cells then faces */
  double *ux, *uy, *ufx, *ufy;         /* u */
  double *u0x, *u0y, *u0fx, *u0fy;     /* u.oldField(0) */
  double *p, *pf;                      /* p */
  double *gpx, *gpy, *gpfx, *gpfy;     /* gradP */
  double *co;
  OrBc ubc[OR_MAX_PATCHES], pbc[OR_MAX_PATCHES];
  OrCrs *uEqn, *pEqn;
  or_solve_cb cb;
  void *cbUser;
  double *xbuf, *bbuf;
  int iters[2];
  double tol;
  int maxIters, precond;
};

static int default_solve(int n, const int *rp, const int *ci, const double *v,
                         const double *b, double *x, void *user) {
  OrFracStep *s = (OrFracStep *)user;
  double rr;
  memset(x, 0, n * sizeof(double));
  int it = or_bicgstab(n, rp, ci, v, b, x, s->tol, s->maxIters, s->precond, &rr);
  return it;
}

OrFracStep *or_fs_create(OrMesh *m, double rho, double mu) {
  OrFracStep *s = (OrFracStep *)xcalloc(1, sizeof(OrFracStep));
  int N = m->nCells, F = m->nFaces;
  s->m = m;
  s->rho = rho;
  s->mu = mu;
#define AL(p, n) s->p = (double *)xcalloc(n, sizeof(double))
  AL(ux, N); AL(uy, N); AL(ufx, F); AL(ufy, F);
  AL(u0x, N); AL(u0y, N); AL(u0fx, F); AL(u0fy, F);
  AL(p, N); AL(pf, F);
  AL(gpx, N); AL(gpy, N); AL(gpfx, F); AL(gpfy, F);
  AL(co, N);
  AL(xbuf, 2 * N); AL(bbuf, 2 * N);
#undef AL
  for (int i = 0; i < OR_MAX_PATCHES; ++i) {
    /* unlisted patches default to NORMAL_GRADIENT (UF/FiniteVolumeField.tpp:104-109) */
    s->ubc[i].type = OR_NORMAL_GRADIENT;
    s->pbc[i].type = OR_NORMAL_GRADIENT;
  }
  s->cb = default_solve;
  s->cbUser = s;
  s->tol = 1e-10;
  s->maxIters = 20000;
  s->precond = 1;
  return s;
}
void or_fs_destroy(OrFracStep *s) {
  if (!s) return;
  double *ps[] = {s->ux, s->uy, s->ufx, s->ufy, s->u0x, s->u0y, s->u0fx,
                  s->u0fy, s->p, s->pf, s->gpx, s->gpy, s->gpfx, s->gpfy,
                  s->co, s->xbuf, s->bbuf};
  for (size_t i = 0; i < sizeof(ps) / sizeof(ps[0]); ++i) free(ps[i]);
  or_crs_destroy(s->uEqn);
  or_crs_destroy(s->pEqn);
  free(s);
}
void or_fs_set_solver(OrFracStep *s, or_solve_cb cb, void *user) {
  if (cb) {
    s->cb = cb;
    s->cbUser = user;
  } else {
    s->cb = default_solve;
    s->cbUser = s;
  }
}

/* setBoundaryTypes/setBoundaryRefValues: UF/FiniteVolumeField.tpp:425-478,
 * UF/VectorFiniteVolumeField.cpp:102-138 -- every face of the patch is set to
 * the reference value. */
int or_fs_set_bc(OrFracStep *s, const char *field, const char *patch, int type,
                 double vx, double vy) {
  OrMesh *m = s->m;
  int id = or_mesh_patch_id(m, patch);
  if (id < 0) return -1;
  if (!strcmp(field, "u")) {
    s->ubc[id].type = type; s->ubc[id].vx = vx; s->ubc[id].vy = vy;
    for (int f = 0; f < m->nFaces; ++f)
      if (m->fPatch[f] == id) { s->ufx[f] = vx; s->ufy[f] = vy; }
  } else if (!strcmp(field, "p")) {
    s->pbc[id].type = type; s->pbc[id].vx = vx;
    for (int f = 0; f < m->nFaces; ++f)
      if (m->fPatch[f] == id) s->pf[f] = vx;
  } else
    return -2;
  return 0;
}

static int bc_type(const OrBc *bc, const OrMesh *m, int f) {
  int p = m->fPatch[f];
  return p < 0 ? OR_NORMAL_GRADIENT : bc[p].type;
}

/* Face::distanceWeight: UG/Face/Face.cpp:60-64 */
static double face_dist_weight(const OrMesh *m, int f) {
  int l = m->fL[f], r = m->fR[f];
  double ax = m->fCx[f] - m->cCx[r], ay = m->fCy[f] - m->cCy[r];
  double bx = m->fCx[f] - m->cCx[l], by = m->fCy[f] - m->cCy[l];
  double l1 = sqrt(ax * ax + ay * ay), l2 = sqrt(bx * bx + by * by);
  return l1 / (l1 + l2);
}

/* VectorFiniteVolumeField::setBoundaryFaces: UF/VectorFiniteVolumeField.cpp:140-161 */
static void u_set_boundary_faces(OrFracStep *s) {
  OrMesh *m = s->m;
  for (int f = 0; f < m->nFaces; ++f) {
    if (m->fR[f] >= 0 || m->fPatch[f] < 0) continue;
    int l = m->fL[f];
    switch (s->ubc[m->fPatch[f]].type) {
    case OR_FIXED:
      break;
    case OR_NORMAL_GRADIENT:
      s->ufx[f] = s->ux[l];
      s->ufy[f] = s->uy[l];
      break;
    case OR_SYMMETRY: {
      double nx = m->fNx[f], ny = m->fNy[f];
      double d = s->ux[l] * nx + s->uy[l] * ny, mm = nx * nx + ny * ny;
      s->ufx[f] = s->ux[l] - d * nx / mm;
      s->ufy[f] = s->uy[l] - d * ny / mm;
    } break;
    }
  }
}
/* FiniteVolumeField<T>::interpolateFaces(DISTANCE): UF/FiniteVolumeField.tpp:129-162 */
static void u_interpolate_faces(OrFracStep *s) {
  OrMesh *m = s->m;
  for (int f = 0; f < m->nFaces; ++f) {
    if (m->fR[f] < 0) continue;
    double g = face_dist_weight(m, f);
    int l = m->fL[f], r = m->fR[f];
    s->ufx[f] = g * s->ux[l] + (1. - g) * s->ux[r];
    s->ufy[f] = g * s->uy[l] + (1. - g) * s->uy[r];
  }
  u_set_boundary_faces(s);
}
/* FiniteVolumeField<Scalar>::setBoundaryFaces: UF/FiniteVolumeField.tpp:164-182 */
static void p_set_boundary_faces(OrFracStep *s) {
  OrMesh *m = s->m;
  for (int f = 0; f < m->nFaces; ++f) {
    if (m->fR[f] >= 0 || m->fPatch[f] < 0) continue;
    int t = s->pbc[m->fPatch[f]].type;
    if (t == OR_NORMAL_GRADIENT || t == OR_SYMMETRY) s->pf[f] = s->p[m->fL[f]];
  }
}
/* ScalarGradient::computeFaces + compute(FACE_TO_CELL): UF/ScalarGradient.cpp:34-74 */
static void grad_p_compute(OrFracStep *s) {
  OrMesh *m = s->m;
  for (int f = 0; f < m->nFaces; ++f) {
    int l = m->fL[f], r = m->fR[f];
    if (r >= 0) {
      double rx = m->cCx[r] - m->cCx[l], ry = m->cCy[r] - m->cCy[l];
      double d = s->p[r] - s->p[l], mm = rx * rx + ry * ry;
      s->gpfx[f] = d * rx / mm;
      s->gpfy[f] = d * ry / mm;
    } else {
      double rx = m->fCx[f] - m->cCx[l], ry = m->fCy[f] - m->cCy[l];
      double d = s->pf[f] - s->p[l], mm = rx * rx + ry * ry;
      s->gpfx[f] = d * rx / mm;
      s->gpfy[f] = d * ry / mm;
    }
  }
  for (int c = 0; c < m->nCells; ++c) {
    if (m->owner[c] != m->rank) continue;
    double sx = 0., sy = 0., tx = 0., ty = 0.;
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      double ax = fabs(m->ilSx[j]), ay = fabs(m->ilSy[j]);
      tx += s->gpfx[m->ilFace[j]] * ax;
      ty += s->gpfy[m->ilFace[j]] * ay;
      sx += ax;
      sy += ay;
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      double ax = fabs(m->blSx[j]), ay = fabs(m->blSy[j]);
      tx += s->gpfx[m->blFace[j]] * ax;
      ty += s->gpfy[m->blFace[j]] * ay;
      sx += ax;
      sy += ay;
    }
    s->gpx[c] = tx / sx;
    s->gpy[c] = ty / sy;
  }
}

/* FractionalStep::initialize: US/FractionalStep.cpp:25-28 */
void or_fs_initialize(OrFracStep *s) {
  u_interpolate_faces(s);
  p_set_boundary_faces(s);
}

/* FiniteVolumeEquation<Vector2D>::add(cell,nb,Scalar): UE/VectorFiniteVolumeEquation.cpp:23-31
 * with IndexMap rows k*nLocal + local, cols offset + k*nLocal (UE/IndexMap.cpp:29-36). */
static void veq_add(OrFracStep *s, OrCrs *e, int cell, int nb, double v) {
  OrMesh *m = s->m;
  int nl = m->nLocal;
  /* vector IndexMap: 2 indices; global = 2*rowOffset + k*nLocal + local for owned
   * cells; ghosts carry their owner's numbering, which needs the owner's nLocal. */
  int r0 = m->localRow[cell];
  int gx = 2 * m->rowOffset + m->localRow[nb];
  or_crs_add_coeff(e, r0, gx, v);
  or_crs_add_coeff(e, nl + r0, gx + nl, v);
}

/* uEqn_ = (fv::ddt(u,dt) + fv::div(u,u,0.) == fv::laplacian(mu/rho,u,0.5) - src::src(gradP)):
 * US/FractionalStep.cpp:79-83; UD/TimeDerivative.h:37-48; UD/Divergence.h:8-53;
 * UD/Laplacian.cpp:5-64; UD/Source.cpp:86-95; algebra M/CrsEquation.cpp:185-311.
 * Single-process numbering only (ghost columns of a vector equation are not
 * needed by any test). */
void or_fs_assemble_u(OrFracStep *s, double dt) {
  OrMesh *m = s->m;
  int N = m->nCells, nl = m->nLocal;
  double theta;
  /* fv::ddt */
  OrCrs *e1 = or_crs_create(2 * nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    veq_add(s, e1, c, c, m->vol[c] / dt);
    int r = m->localRow[c];
    e1->rhs[r] += -m->vol[c] * s->u0x[c] / dt;
    e1->rhs[nl + r] += -m->vol[c] * s->u0y[c] / dt;
  }
  /* fv::div(u,u,theta=0) */
  theta = 0.;
  OrCrs *e2 = or_crs_create(2 * nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      int f = m->ilFace[j], nb = m->ilCell[j];
      double flux = s->ufx[f] * m->ilSx[j] + s->ufy[f] * m->ilSy[j];
      double flux0 = s->u0fx[f] * m->ilSx[j] + s->u0fy[f] * m->ilSy[j];
      veq_add(s, e2, c, c, theta * fmax(flux, 0.));
      veq_add(s, e2, c, nb, theta * fmin(flux, 0.));
      double a = (1. - theta) * fmax(flux0, 0.);
      e2->rhs[r] += a * s->u0x[c];
      e2->rhs[nl + r] += a * s->u0y[c];
      double b = (1. - theta) * fmin(flux0, 0.);
      e2->rhs[r] += b * s->u0x[nb];
      e2->rhs[nl + r] += b * s->u0y[nb];
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      int f = m->blFace[j];
      double flux = s->ufx[f] * m->blSx[j] + s->ufy[f] * m->blSy[j];
      double flux0 = s->u0fx[f] * m->blSx[j] + s->u0fy[f] * m->blSy[j];
      switch (bc_type(s->ubc, m, f)) {
      case OR_FIXED:
        e2->rhs[r] += theta * flux * s->ufx[f];
        e2->rhs[nl + r] += theta * flux * s->ufy[f];
        e2->rhs[r] += (1. - theta) * flux0 * s->u0fx[f];
        e2->rhs[nl + r] += (1. - theta) * flux0 * s->u0fy[f];
        break;
      case OR_NORMAL_GRADIENT:
        veq_add(s, e2, c, c, theta * flux);
        e2->rhs[r] += (1. - theta) * flux0 * s->u0x[c];
        e2->rhs[nl + r] += (1. - theta) * flux0 * s->u0y[c];
        break;
      default:
        break;
      }
    }
  }
  or_crs_add_eq(e1, e2); /* operator+ */
  or_crs_destroy(e2);
  /* fv::laplacian(mu/rho, u, 0.5) */
  theta = 0.5;
  double gamma = s->mu / s->rho;
  OrCrs *e3 = or_crs_create(2 * nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      int nb = m->ilCell[j];
      double coeff = gamma * (m->ilRcx[j] * m->ilSx[j] + m->ilRcy[j] * m->ilSy[j]) /
                     (m->ilRcx[j] * m->ilRcx[j] + m->ilRcy[j] * m->ilRcy[j]);
      veq_add(s, e3, c, nb, theta * coeff);
      veq_add(s, e3, c, c, theta * -coeff);
      double a = (1. - theta) * coeff;
      e3->rhs[r] += a * (s->u0x[nb] - s->u0x[c]);
      e3->rhs[nl + r] += a * (s->u0y[nb] - s->u0y[c]);
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      int f = m->blFace[j];
      double coeff = gamma * (m->blRfx[j] * m->blSx[j] + m->blRfy[j] * m->blSy[j]) /
                     (m->blRfx[j] * m->blRfx[j] + m->blRfy[j] * m->blRfy[j]);
      switch (bc_type(s->ubc, m, f)) {
      case OR_FIXED: {
        veq_add(s, e3, c, c, theta * -coeff);
        e3->rhs[r] += theta * coeff * s->ufx[f];
        e3->rhs[nl + r] += theta * coeff * s->ufy[f];
        double a = (1. - theta) * coeff;
        e3->rhs[r] += a * (s->u0fx[f] - s->u0x[c]);
        e3->rhs[nl + r] += a * (s->u0fy[f] - s->u0y[c]);
      } break;
      default: /* NORMAL_GRADIENT; SYMMETRY tensor term not restated (no config uses it) */
        break;
      }
    }
  }
  /* - src::src(gradP) */
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    e3->rhs[r] -= s->gpx[c] * m->vol[c];
    e3->rhs[nl + r] -= s->gpy[c] * m->vol[c];
  }
  or_crs_sub_eq(e1, e3); /* operator== */
  or_crs_destroy(e3);
  or_crs_destroy(s->uEqn);
  s->uEqn = e1;
}

/* pEqn_ = (fv::laplacian(dt, p) == src::div(u)): US/FractionalStep.cpp:96-97;
 * UD/Laplacian.h:49-83 (nb inserted before the diagonal); UD/Source.cpp:5-25. */
static void src_div_u(OrFracStep *s, double *out) {
  OrMesh *m = s->m;
  for (int c = 0; c < m->nCells; ++c) {
    if (m->owner[c] != m->rank) continue;
    double d = 0.;
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j)
      d += s->ufx[m->ilFace[j]] * m->ilSx[j] + s->ufy[m->ilFace[j]] * m->ilSy[j];
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j)
      d += s->ufx[m->blFace[j]] * m->blSx[j] + s->ufy[m->blFace[j]] * m->blSy[j];
    out[m->localRow[c]] = d;
  }
}

static OrCrs *laplacian_p(OrFracStep *s, double gammaConst,
                          const double *gammaFace, int diagFirst) {
  OrMesh *m = s->m;
  OrCrs *e = or_crs_create(m->nLocal, 5);
  for (int c = 0; c < m->nCells; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      double g = gammaFace ? gammaFace[m->ilFace[j]] : gammaConst;
      double coeff = g * (m->ilRcx[j] * m->ilSx[j] + m->ilRcy[j] * m->ilSy[j]) /
                     (m->ilRcx[j] * m->ilRcx[j] + m->ilRcy[j] * m->ilRcy[j]);
      int gcol = m->globalRow[m->ilCell[j]], gdiag = m->globalRow[c];
      if (diagFirst) { /* UD/Laplacian.h:141-142 */
        or_crs_add_coeff(e, r, gdiag, -coeff);
        or_crs_add_coeff(e, r, gcol, coeff);
      } else { /* UD/Laplacian.h:57-58 */
        or_crs_add_coeff(e, r, gcol, coeff);
        or_crs_add_coeff(e, r, gdiag, -coeff);
      }
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      int f = m->blFace[j];
      double g = gammaFace ? gammaFace[f] : gammaConst;
      double coeff = g * (m->blRfx[j] * m->blSx[j] + m->blRfy[j] * m->blSy[j]) /
                     (m->blRfx[j] * m->blRfx[j] + m->blRfy[j] * m->blRfy[j]);
      if (bc_type(s->pbc, m, f) == OR_FIXED) {
        or_crs_add_coeff(e, r, m->globalRow[c], -coeff);
        e->rhs[r] += coeff * s->pf[f];
      }
    }
  }
  return e;
}

void or_fs_assemble_p(OrFracStep *s, double dt) {
  OrCrs *e = laplacian_p(s, dt, NULL, 0);
  src_div_u(s, s->bbuf);
  or_crs_sub_vec(e, s->bbuf); /* operator==(Vector) */
  or_crs_destroy(s->pEqn);
  s->pEqn = e;
}

/* FractionalStepMultiphase::solvePEqn: US/FractionalStepMultiphase.cpp:153;
 * UD/Laplacian.h:132-167 (diagonal inserted first). */
OrCrs *or_op_laplacian_field(OrFracStep *s, const double *gammaFace) {
  OrCrs *e = laplacian_p(s, 0., gammaFace, 1);
  src_div_u(s, s->bbuf);
  or_crs_sub_vec(e, s->bbuf);
  return e;
}


/* ---------------------------------------------- multiphase-style operators */
/* FiniteVolumeEquation<Vector2D> operator*(ScalarField rho, eqn): scaleRow of
 * rows (P,0),(P,1): UE/VectorFiniteVolumeEquation.cpp:163-170, M/CrsEquation.cpp:153-159 */
static void veq_scale_rows(OrFracStep *s, OrCrs *e, const double *rho) {
  OrMesh *m = s->m;
  for (int c = 0; c < m->nCells; ++c) {
    if (m->owner[c] != m->rank) continue;
    or_crs_scale_row(e, m->localRow[c], rho[c]);
    or_crs_scale_row(e, m->nLocal + m->localRow[c], rho[c]);
  }
}

/* uEqn_ = (rho*fv::ddt(u,dt) + rho*fv::dive(u,u,0.5) == fv::laplacian(mu,u,0.5) + src::src(f)):
 * US/FractionalStepMultiphase.cpp:111-112; UD/TimeDerivative.h:37-48; UD/ExplicitDivergence.h:7-51
 * (oldField(0) and oldField(1) are the SAME buffer); UD/Laplacian.cpp:66-119
 * (field gamma, diagonal added first); UD/Source.cpp:86-95.  f = cell vector field (fx, fy). */
OrCrs *or_op_ueqn_multiphase(OrFracStep *s, double dt, const double *rhoCell,
                             const double *muFace, const double *mu0Face,
                             const double *fx, const double *fy) {
  OrMesh *m = s->m;
  int N = m->nCells, nl = m->nLocal;
  double theta = 0.5;
  OrCrs *e1 = or_crs_create(2 * nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    veq_add(s, e1, c, c, m->vol[c] / dt);
    int r = m->localRow[c];
    e1->rhs[r] += -m->vol[c] * s->u0x[c] / dt;
    e1->rhs[nl + r] += -m->vol[c] * s->u0y[c] / dt;
  }
  veq_scale_rows(s, e1, rhoCell);
  OrCrs *e2 = or_crs_create(2 * nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      int f = m->ilFace[j], nb = m->ilCell[j];
      double flux0 = s->u0fx[f] * m->ilSx[j] + s->u0fy[f] * m->ilSy[j];
      double flux1 = flux0; /* aliased history */
      double a0 = theta * fmax(flux0, 0.), b0 = theta * fmin(flux0, 0.);
      double a1 = (1. - theta) * fmax(flux1, 0.), b1 = (1. - theta) * fmin(flux1, 0.);
      e2->rhs[r] += a0 * s->u0x[c]; e2->rhs[nl + r] += a0 * s->u0y[c];
      e2->rhs[r] += b0 * s->u0x[nb]; e2->rhs[nl + r] += b0 * s->u0y[nb];
      e2->rhs[r] += a1 * s->u0x[c]; e2->rhs[nl + r] += a1 * s->u0y[c];
      e2->rhs[r] += b1 * s->u0x[nb]; e2->rhs[nl + r] += b1 * s->u0y[nb];
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      int f = m->blFace[j];
      int t = bc_type(s->ubc, m, f);
      if (t != OR_FIXED && t != OR_NORMAL_GRADIENT) continue;
      double flux0 = s->u0fx[f] * m->blSx[j] + s->u0fy[f] * m->blSy[j];
      double flux1 = flux0;
      e2->rhs[r] += theta * flux0 * s->u0fx[f]; e2->rhs[nl + r] += theta * flux0 * s->u0fy[f];
      e2->rhs[r] += (1. - theta) * flux1 * s->u0fx[f]; e2->rhs[nl + r] += (1. - theta) * flux1 * s->u0fy[f];
    }
  }
  veq_scale_rows(s, e2, rhoCell);
  or_crs_add_eq(e1, e2);
  or_crs_destroy(e2);
  OrCrs *e3 = or_crs_create(2 * nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      int f = m->ilFace[j], nb = m->ilCell[j];
      double g = (m->ilRcx[j] * m->ilSx[j] + m->ilRcy[j] * m->ilSy[j]) /
                 (m->ilRcx[j] * m->ilRcx[j] + m->ilRcy[j] * m->ilRcy[j]);
      double coeff = muFace[f] * g, coeff0 = mu0Face[f] * g;
      veq_add(s, e3, c, c, theta * -coeff);
      veq_add(s, e3, c, nb, theta * coeff);
      double a = (1. - theta) * coeff0;
      e3->rhs[r] += a * (s->u0x[nb] - s->u0x[c]);
      e3->rhs[nl + r] += a * (s->u0y[nb] - s->u0y[c]);
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      int f = m->blFace[j];
      if (bc_type(s->ubc, m, f) != OR_FIXED) continue;
      double g = (m->blRfx[j] * m->blSx[j] + m->blRfy[j] * m->blSy[j]) /
                 (m->blRfx[j] * m->blRfx[j] + m->blRfy[j] * m->blRfy[j]);
      double coeff = muFace[f] * g, coeff0 = mu0Face[f] * g;
      veq_add(s, e3, c, c, theta * -coeff);
      e3->rhs[r] += theta * coeff * s->ufx[f]; e3->rhs[nl + r] += theta * coeff * s->ufy[f];
      double a = (1. - theta) * coeff0;
      e3->rhs[r] += a * (s->u0fx[f] - s->u0x[c]);
      e3->rhs[nl + r] += a * (s->u0fy[f] - s->u0y[c]);
    }
  }
  for (int c = 0; c < N; ++c) { /* + src::src(f) */
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    e3->rhs[r] += fx[c] * m->vol[c];
    e3->rhs[nl + r] += fy[c] * m->vol[c];
  }
  or_crs_sub_eq(e1, e3);
  or_crs_destroy(e3);
  return e1;
}

/* scalar transport on the "p" field: (fv::ddt(rho, phi, dt) + fv::div(u, phi, theta) == 0):
 * UD/TimeDerivative.h:21-35 (rho(cell), rho0(cell)); UD/Divergence.h:8-53 with the
 * theta-weighted matrix part and FIXED / NORMAL_GRADIENT boundary handling.
 * phi = s->p / s->pf, old level phi0 given. */
OrCrs *or_op_scalar_transport(OrFracStep *s, double dt, double theta, const double *rho,
                              const double *rho0, const double *phi0, const double *phi0f) {
  OrMesh *m = s->m;
  int N = m->nCells, nl = m->nLocal;
  OrCrs *e1 = or_crs_create(nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    or_crs_add_coeff(e1, r, m->globalRow[c], rho[c] * m->vol[c] / dt);
    e1->rhs[r] += -rho0[c] * m->vol[c] * phi0[c] / dt;
  }
  OrCrs *e2 = or_crs_create(nl, 5);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      int f = m->ilFace[j], nb = m->ilCell[j];
      double flux = s->ufx[f] * m->ilSx[j] + s->ufy[f] * m->ilSy[j];
      double flux0 = s->u0fx[f] * m->ilSx[j] + s->u0fy[f] * m->ilSy[j];
      or_crs_add_coeff(e2, r, m->globalRow[c], theta * fmax(flux, 0.));
      or_crs_add_coeff(e2, r, m->globalRow[nb], theta * fmin(flux, 0.));
      e2->rhs[r] += (1. - theta) * fmax(flux0, 0.) * phi0[c];
      e2->rhs[r] += (1. - theta) * fmin(flux0, 0.) * phi0[nb];
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      int f = m->blFace[j];
      double flux = s->ufx[f] * m->blSx[j] + s->ufy[f] * m->blSy[j];
      double flux0 = s->u0fx[f] * m->blSx[j] + s->u0fy[f] * m->blSy[j];
      switch (bc_type(s->pbc, m, f)) {
      case OR_FIXED:
        e2->rhs[r] += theta * flux * s->pf[f];
        e2->rhs[r] += (1. - theta) * flux0 * phi0f[f];
        break;
      case OR_NORMAL_GRADIENT:
        or_crs_add_coeff(e2, r, m->globalRow[c], theta * flux);
        e2->rhs[r] += (1. - theta) * flux0 * phi0[c];
        break;
      default:
        break;
      }
    }
  }
  or_crs_add_eq(e1, e2);
  or_crs_destroy(e2);
  return e1;
}


/* --------------------------------------------------------------- CICSAM (A9) */
static double clampd(double v, double lo, double hi) { return fmax(fmin(v, hi), lo); }
/* cicsam::hc / cicsam::uq: UD/Cicsam.cpp:5-17 */
static double cicsam_hc(double g, double co) { return g >= 0 && g <= 1 ? fmin(1., g / co) : g; }
static double cicsam_uq(double g, double co) {
  return g >= 0 && g <= 1 ? fmin((8. * co * g + (1. - co) * (6. * g + 3.)) / 8., cicsam_hc(g, co)) : g;
}
/* cicsam::faceInterpolationWeights: UD/Cicsam.cpp:19-66.  gamma = the scalar field "p" (cells),
 * gradGamma = (gx, gy) cell values, u = the vector field's faces. */
void or_cicsam_weights(OrFracStep *s, double dt, const double *gx, const double *gy, double *beta) {
  OrMesh *m = s->m;
  const double k = 1.;
  for (int f = 0; f < m->nFaces; ++f) beta[f] = 0.;
  for (int f = 0; f < m->nFaces; ++f) {
    int l = m->fL[f], r = m->fR[f];
    if (r < 0) continue;
    double sx, sy;
    outward_norm(m, f, m->cCx[l], m->cCy[l], &sx, &sy);
    double flux = s->ufx[f] * sx + s->ufy[f] * sy;
    int d = flux > 0. ? l : r, a = flux <= 0. ? l : r;
    double rcx = m->cCx[a] - m->cCx[d], rcy = m->cCy[a] - m->cCy[d];
    double gD = clampd(s->p[d], 0., 1.), gA = clampd(s->p[a], 0., 1.);
    double gU = clampd(gA - 2. * (rcx * gx[d] + rcy * gy[d]), 0., 1.);
    double gDT = (gD - gU) / (gA - gU);
    if (!isfinite(gDT)) gDT = 0.;
    double coD = 0.;
    for (int j = m->ilPtr[d]; j < m->ilPtr[d + 1]; ++j)
      coD += fmax((s->ufx[m->ilFace[j]] * m->ilSx[j] + s->ufy[m->ilFace[j]] * m->ilSy[j]) / m->vol[d] * dt, 0.);
    for (int j = m->blPtr[d]; j < m->blPtr[d + 1]; ++j)
      coD += fmax((s->ufx[m->blFace[j]] * m->blSx[j] + s->ufy[m->blFace[j]] * m->blSy[j]) / m->vol[d] * dt, 0.);
    double gm = sqrt(gx[d] * gx[d] + gy[d] * gy[d]), rm = sqrt(rcx * rcx + rcy * rcy);
    double thetaF = acos(fabs((gx[d] / gm) * (rcx / rm) + (gy[d] / gm) * (rcy / rm)));
    double psiF = fmin(k * (cos(2 * thetaF) + 1.) / 2., 1.);
    double gFT = psiF * cicsam_hc(gDT, coD) + (1. - psiF) * cicsam_uq(gDT, coD);
    double b = (gFT - gDT) / (1. - gDT);
    if (isfinite(b)) b = fmax(fmin(1., b), 0.);
    else b = 0.;
    beta[f] = b;
  }
}
/* cicsam::div(u, gamma, beta, theta): UD/Cicsam.cpp:89-138 on the scalar field "p"; gamma0 given */
OrCrs *or_op_cicsam_div(OrFracStep *s, double theta, const double *beta, const double *g0, const double *g0f) {
  OrMesh *m = s->m;
  OrCrs *e = or_crs_create(m->nLocal, 5);
  for (int c = 0; c < m->nCells; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j) {
      int f = m->ilFace[j], nb = m->ilCell[j];
      double flux = s->ufx[f] * m->ilSx[j] + s->ufy[f] * m->ilSy[j];
      int donor = flux > 0. ? c : nb, acceptor = flux <= 0. ? c : nb;
      double b = beta[f];
      or_crs_add_coeff(e, r, m->globalRow[donor], theta * (1. - b) * flux);
      or_crs_add_coeff(e, r, m->globalRow[acceptor], theta * b * flux);
      double gF = (1. - b) * g0[donor] + b * g0[acceptor];
      e->rhs[r] += (1. - theta) * flux * gF;
    }
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j) {
      int f = m->blFace[j];
      double flux = s->ufx[f] * m->blSx[j] + s->ufy[f] * m->blSy[j];
      switch (bc_type(s->pbc, m, f)) {
      case OR_FIXED:
        e->rhs[r] += theta * flux * s->pf[f];
        e->rhs[r] += (1. - theta) * flux * g0f[f];
        break;
      case OR_NORMAL_GRADIENT:
        or_crs_add_coeff(e, r, m->globalRow[c], theta * flux);
        e->rhs[r] += (1. - theta) * flux * g0[c];
        break;
      default:
        break;
      }
    }
  }
  return e;
}
/* cicsam::computeMomentumFlux: UD/Cicsam.cpp:69-87 -> rhoU faces (x, y) */
void or_cicsam_momentum_flux(OrFracStep *s, double rho1, double rho2, const double *beta, double *outx, double *outy) {
  OrMesh *m = s->m;
  for (int f = 0; f < m->nFaces; ++f) {
    int l = m->fL[f], r = m->fR[f];
    double g;
    if (r >= 0) {
      double sx, sy;
      outward_norm(m, f, m->cCx[l], m->cCy[l], &sx, &sy);
      double flux = s->ufx[f] * sx + s->ufy[f] * sy;
      int d = flux > 0. ? l : r, a = flux <= 0. ? l : r;
      g = (1. - beta[f]) * s->p[d] + beta[f] * s->p[a];
    } else
      g = s->pf[f];
    double rho = rho1 + clampd(g, 0., 1.) * (rho2 - rho1);
    outx[f] = rho * s->ufx[f];
    outy[f] = rho * s->ufy[f];
  }
}

const OrCrs *or_fs_ueqn(const OrFracStep *s) { return s->uEqn; }
const OrCrs *or_fs_peqn(const OrFracStep *s) { return s->pEqn; }

/* FiniteVolumeEquation<T>::solve: UE/FiniteVolumeEquation.tpp:64-86 --
 * set(rowPtr,colInd,vals); setRhs(-rhs_); solve; mapFromSparseSolver. */
static int eqn_solve(OrFracStep *s, OrCrs *e, double *x) {
  int n = e->nRows;
  for (int i = 0; i < n; ++i) s->bbuf[i] = -e->rhs[i];
  return s->cb(n, e->rowPtr, e->colInd, e->vals, s->bbuf, x, s->cbUser);
}

/* FractionalStep::solve: US/FractionalStep.cpp:36-46, 79-117 */
int or_fs_step(OrFracStep *s, double dt) {
  OrMesh *m = s->m;
  int N = m->nCells, F = m->nFaces, nl = m->nLocal;
  /* solveUEqn: savePreviousTimeStep(dt,1) -- deep copy (UF/FiniteVolumeField.tpp:208-227) */
  memcpy(s->u0x, s->ux, N * sizeof(double));
  memcpy(s->u0y, s->uy, N * sizeof(double));
  memcpy(s->u0fx, s->ufx, F * sizeof(double));
  memcpy(s->u0fy, s->ufy, F * sizeof(double));
  or_fs_assemble_u(s, dt);
  s->iters[0] = eqn_solve(s, s->uEqn, s->xbuf);
  for (int c = 0; c < N; ++c) {
    if (m->owner[c] != m->rank) continue;
    int r = m->localRow[c];
    s->ux[c] = s->xbuf[r];
    s->uy[c] = s->xbuf[nl + r];
  }
  for (int c = 0; c < N; ++c) { /* u += dt*gradP (:87-88) */
    s->ux[c] += dt * s->gpx[c];
    s->uy[c] += dt * s->gpy[c];
  }
  u_interpolate_faces(s);
  /* solvePEqn */
  or_fs_assemble_p(s, dt);
  s->iters[1] = eqn_solve(s, s->pEqn, s->xbuf);
  for (int c = 0; c < N; ++c)
    if (m->owner[c] == m->rank) s->p[c] = s->xbuf[m->localRow[c]];
  p_set_boundary_faces(s);
  grad_p_compute(s);
  /* correctVelocity (:109-117) */
  for (int c = 0; c < N; ++c) {
    s->ux[c] -= dt * s->gpx[c];
    s->uy[c] -= dt * s->gpy[c];
  }
  for (int f = 0; f < F; ++f) {
    s->ufx[f] -= dt * s->gpfx[f];
    s->ufy[f] -= dt * s->gpfy[f];
  }
  return 0;
}

/* FractionalStep::maxDivergenceError: US/FractionalStep.cpp:119-135 (returns
 * the max of |div|; the reference stores the signed value, Appendix A) */
double or_fs_max_divergence(const OrFracStep *s) {
  const OrMesh *m = s->m;
  double mx = 0.;
  for (int c = 0; c < m->nCells; ++c) {
    double d = 0.;
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j)
      d += s->ufx[m->ilFace[j]] * m->ilSx[j] + s->ufy[m->ilFace[j]] * m->ilSy[j];
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j)
      d += s->ufx[m->blFace[j]] * m->blSx[j] + s->ufy[m->blFace[j]] * m->blSy[j];
    if (fabs(d) > mx) mx = fabs(d);
  }
  return mx;
}
/* FractionalStep::maxCourantNumber: US/FractionalStep.cpp:48-66 */
double or_fs_max_courant(OrFracStep *s, double dt) {
  const OrMesh *m = s->m;
  double mx = 0.;
  for (int c = 0; c < m->nCells; ++c) {
    double co = 0.;
    for (int j = m->ilPtr[c]; j < m->ilPtr[c + 1]; ++j)
      co += fmax(s->ufx[m->ilFace[j]] * m->ilSx[j] + s->ufy[m->ilFace[j]] * m->ilSy[j], 0.);
    for (int j = m->blPtr[c]; j < m->blPtr[c + 1]; ++j)
      co += fmax(s->ufx[m->blFace[j]] * m->blSx[j] + s->ufy[m->blFace[j]] * m->blSy[j], 0.);
    co *= dt / m->vol[c];
    s->co[c] = co;
    if (co > mx) mx = co;
  }
  return mx;
}
int or_fs_last_iters(const OrFracStep *s, int which) { return s->iters[which & 1]; }

long or_fs_array(OrFracStep *s, const char *name, double **ptr) {
  int N = s->m->nCells, F = s->m->nFaces;
#define FA(nm, p, n) if (!strcmp(name, nm)) { *ptr = s->p; return n; }
  FA("ux", ux, N) FA("uy", uy, N) FA("ufx", ufx, F) FA("ufy", ufy, F)
  FA("u0x", u0x, N) FA("u0y", u0y, N) FA("u0fx", u0fx, F) FA("u0fy", u0fy, F)
  FA("p", p, N) FA("pf", pf, F)
  FA("gpx", gpx, N) FA("gpy", gpy, N) FA("gpfx", gpfx, F) FA("gpfy", gpfy, F)
  FA("co", co, N)
#undef FA
  return -1;
}
void or_fs_set_solver_params(OrFracStep *s, double tol, int maxIters,
                             int precond) {
  s->tol = tol;
  s->maxIters = maxIters;
  s->precond = precond;
}

/* ------------------------------------------------- CPU BiCGStab (C2 baseline) */
/* Right-preconditioned BiCGStab as in Belos/Eigen (S4: the algorithm the README
 * and the Belos default name), OpenMP over rows.  -1 padded columns are skipped
 * (M/EigenSparseMatrixSolver.cpp:33-36). */
int or_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* launchers such as torch.distributed.run export OMP_NUM_THREADS=1: the timed CPU arm sets its thread count itself */
int or_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

static void spmv(int n, const int *rp, const int *ci, const double *v,
                 const double *x, double *y) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    double a = 0.;
    for (int j = rp[i]; j < rp[i + 1]; ++j)
      if (ci[j] >= 0) a += v[j] * x[ci[j]];
    y[i] = a;
  }
}
static double dotp(int n, const double *a, const double *b) {
  double s = 0.;
#pragma omp parallel for reduction(+ : s) schedule(static)
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

typedef struct {
  int n;
  const int *rp, *ci;
  double *lu;   /* ILU(0) factors on the pattern of A */
  int *diag;    /* slot of the diagonal per row */
  double *dinv; /* Jacobi */
  int kind;
  int nBlocks;        /* > 0: rows are grouped in independent sets (multicolour ordering): */
  const int *blockPtr; /* the sweeps run set by set, OpenMP-parallel inside a set */
} Precond;

static void precond_setup(Precond *P, int n, const int *rp, const int *ci,
                          const double *v, int kind) {
  P->n = n; P->rp = rp; P->ci = ci; P->kind = kind;
  P->nBlocks = 0; P->blockPtr = NULL;
  P->lu = NULL; P->diag = NULL; P->dinv = NULL;
  if (kind == 1) {
    P->dinv = (double *)xcalloc(n, sizeof(double));
    for (int i = 0; i < n; ++i) {
      double d = 1.;
      for (int j = rp[i]; j < rp[i + 1]; ++j)
        if (ci[j] == i) d = v[j];
      P->dinv[i] = 1. / d;
    }
  } else if (kind == 2) {
    /* ILU(0), IKJ variant on unsorted rows; sequential */
    int nnz = rp[n], maxLen = 1;
    P->lu = (double *)xcalloc(nnz, sizeof(double));
    memcpy(P->lu, v, nnz * sizeof(double));
    P->diag = (int *)xcalloc(n, sizeof(int));
    int *pos = (int *)xcalloc(n, sizeof(int));
    for (int i = 0; i < n; ++i) {
      pos[i] = -1;
      if (rp[i + 1] - rp[i] > maxLen) maxLen = rp[i + 1] - rp[i];
      for (int j = rp[i]; j < rp[i + 1]; ++j)
        if (ci[j] == i) P->diag[i] = j;
    }
    int *lo = (int *)xcalloc(maxLen, sizeof(int));
    for (int i = 0; i < n; ++i) {
      int nlo = 0;
      for (int j = rp[i]; j < rp[i + 1]; ++j) {
        int k = ci[j];
        if (k < 0) continue;
        pos[k] = j;
        if (k < i) { /* insertion sort of lower slots by column */
          int q = nlo++;
          while (q > 0 && ci[lo[q - 1]] > k) {
            lo[q] = lo[q - 1];
            --q;
          }
          lo[q] = j;
        }
      }
      for (int a = 0; a < nlo; ++a) {
        int j = lo[a], k = ci[j];
        double lik = P->lu[j] / P->lu[P->diag[k]];
        P->lu[j] = lik;
        for (int q = rp[k]; q < rp[k + 1]; ++q) {
          int c = ci[q];
          if (c > k && pos[c] >= 0) P->lu[pos[c]] -= lik * P->lu[q];
        }
      }
      for (int j = rp[i]; j < rp[i + 1]; ++j)
        if (ci[j] >= 0) pos[ci[j]] = -1;
    }
    free(lo);
    free(pos);
  }
}
static void precond_apply(const Precond *P, const double *r, double *z) {
  int n = P->n;
  if (P->kind == 0) {
    memcpy(z, r, n * sizeof(double));
  } else if (P->kind == 1) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) z[i] = r[i] * P->dinv[i];
  } else if (P->nBlocks > 0) {
    /* multicolour ordering: rows of a set are independent -> parallel sweeps */
    for (int b = 0; b < P->nBlocks; ++b) {
#pragma omp parallel for schedule(static)
      for (int i = P->blockPtr[b]; i < P->blockPtr[b + 1]; ++i) {
        double a = r[i];
        for (int j = P->rp[i]; j < P->rp[i + 1]; ++j)
          if (P->ci[j] >= 0 && P->ci[j] < i) a -= P->lu[j] * z[P->ci[j]];
        z[i] = a;
      }
    }
    for (int b = P->nBlocks - 1; b >= 0; --b) {
#pragma omp parallel for schedule(static)
      for (int i = P->blockPtr[b]; i < P->blockPtr[b + 1]; ++i) {
        double a = z[i];
        for (int j = P->rp[i]; j < P->rp[i + 1]; ++j)
          if (P->ci[j] > i) a -= P->lu[j] * z[P->ci[j]];
        z[i] = a / P->lu[P->diag[i]];
      }
    }
  } else {
    for (int i = 0; i < n; ++i) {
      double a = r[i];
      for (int j = P->rp[i]; j < P->rp[i + 1]; ++j)
        if (P->ci[j] >= 0 && P->ci[j] < i) a -= P->lu[j] * z[P->ci[j]];
      z[i] = a;
    }
    for (int i = n - 1; i >= 0; --i) {
      double a = z[i];
      for (int j = P->rp[i]; j < P->rp[i + 1]; ++j)
        if (P->ci[j] > i) a -= P->lu[j] * z[P->ci[j]];
      z[i] = a / P->lu[P->diag[i]];
    }
  }
}
static void precond_free(Precond *P) {
  free(P->lu);
  free(P->diag);
  free(P->dinv);
}

/* greedy multicolouring in the given order (same rule as the CUDA path's symbolic
 * phase): colour of row i = smallest colour not used by its neighbours j < i.
 * new2old = rows sorted by colour (stable), blockPtr[nColours+1]; returns nColours. */
int or_multicolor_order(int n, const int *rp, const int *ci, int *new2old, int *blockPtr, int maxColours) {
  int *col = (int *)xcalloc(n, sizeof(int));
  int nc = 1;
  for (int i = 0; i < n; ++i) {
    unsigned long long used = 0ull;
    for (int j = rp[i]; j < rp[i + 1]; ++j)
      if (ci[j] >= 0 && ci[j] < i && col[ci[j]] < 64) used |= 1ull << col[ci[j]];
    int c = 0;
    while ((used >> c) & 1ull) ++c;
    col[i] = c;
    if (c + 1 > nc) nc = c + 1;
  }
  if (nc > maxColours) { free(col); return -nc; }
  for (int c = 0; c <= nc; ++c) blockPtr[c] = 0;
  for (int i = 0; i < n; ++i) blockPtr[col[i] + 1]++;
  for (int c = 0; c < nc; ++c) blockPtr[c + 1] += blockPtr[c];
  int *fill = (int *)xcalloc(nc, sizeof(int));
  for (int i = 0; i < n; ++i) new2old[blockPtr[col[i]] + fill[col[i]]++] = i;
  free(fill);
  free(col);
  return nc;
}

static int bicgstab_impl(int n, const int *rp, const int *ci, const double *v,
                const double *b, double *x, double tol, int maxIters,
                int precond, double *relres, int nBlocks, const int *blockPtr);

int or_bicgstab(int n, const int *rp, const int *ci, const double *v,
                const double *b, double *x, double tol, int maxIters,
                int precond, double *relres) {
  return bicgstab_impl(n, rp, ci, v, b, x, tol, maxIters, precond, relres, 0, NULL);
}
/* ILU(0) on a system whose rows are already grouped in independent sets (blockPtr) */
int or_bicgstab_blocks(int n, const int *rp, const int *ci, const double *v,
                       const double *b, double *x, double tol, int maxIters,
                       int nBlocks, const int *blockPtr, double *relres) {
  return bicgstab_impl(n, rp, ci, v, b, x, tol, maxIters, 2, relres, nBlocks, blockPtr);
}

static int bicgstab_impl(int n, const int *rp, const int *ci, const double *v,
                const double *b, double *x, double tol, int maxIters,
                int precond, double *relres, int nBlocks, const int *blockPtr) {
  double *r = (double *)xcalloc(n, sizeof(double));
  double *r0 = (double *)xcalloc(n, sizeof(double));
  double *p = (double *)xcalloc(n, sizeof(double));
  double *vv = (double *)xcalloc(n, sizeof(double));
  double *s = (double *)xcalloc(n, sizeof(double));
  double *t = (double *)xcalloc(n, sizeof(double));
  double *ph = (double *)xcalloc(n, sizeof(double));
  double *sh = (double *)xcalloc(n, sizeof(double));
  Precond P;
  precond_setup(&P, n, rp, ci, v, precond);
  P.nBlocks = nBlocks;
  P.blockPtr = blockPtr;
  spmv(n, rp, ci, v, x, r);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < n; ++i) {
    r[i] = b[i] - r[i];
    r0[i] = r[i];
  }
  double bnorm = sqrt(dotp(n, b, b));
  if (bnorm == 0.) bnorm = 1.;
  double rho = 1., alpha = 1., omega = 1.;
  double rn = sqrt(dotp(n, r, r));
  int it = 0;
  while (it < maxIters && rn / bnorm > tol) {
    double rhoNew = dotp(n, r0, r);
    if (rhoNew == 0.) break;
    if (it == 0) {
      memcpy(p, r, n * sizeof(double));
    } else {
      double beta = (rhoNew / rho) * (alpha / omega);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i) p[i] = r[i] + beta * (p[i] - omega * vv[i]);
    }
    rho = rhoNew;
    precond_apply(&P, p, ph);
    spmv(n, rp, ci, v, ph, vv);
    alpha = rho / dotp(n, r0, vv);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) s[i] = r[i] - alpha * vv[i];
    precond_apply(&P, s, sh);
    spmv(n, rp, ci, v, sh, t);
    double tt = dotp(n, t, t);
    omega = tt == 0. ? 0. : dotp(n, t, s) / tt;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n; ++i) {
      x[i] += alpha * ph[i] + omega * sh[i];
      r[i] = s[i] - omega * t[i];
    }
    rn = sqrt(dotp(n, r, r));
    ++it;
    if (omega == 0.) break;
  }
  if (relres) *relres = rn / bnorm;
  precond_free(&P);
  free(r); free(r0); free(p); free(vv); free(s); free(t); free(ph); free(sh);
  return it;
}
