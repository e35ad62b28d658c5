// fv.cu -- device-resident fields and equations (Seam 2): the fv:: / src::
// operators as cell-parallel gather kernels on the canonical pattern (K1-K6),
// field glue (K14) and FiniteVolumeEquation<T>::solve without host round trips.
//
// No atomics anywhere: every row is written by exactly one lane, which gathers
// the per-face factors (g_f, S_f, weights) precomputed once per mesh; the slot of
// a link in its row is fixed by the mesh (entry k+1 of the row = interior link k,
// the face->slot map exported as slotL/slotR).  Boundary links are handled by a
// second, tiny launch over the boundary cells.
//
// Reference operators (under /root/reference/src/2D/Unstructured/FiniteVolume):
//   Discretization/TimeDerivative.h:7-48   Discretization/Divergence.h:8-53
//   Discretization/ExplicitDivergence.h:7-51   Discretization/Laplacian.h:7-167
//   Discretization/Laplacian.cpp:5-119     Discretization/Source.cpp:5-25,77-95
//   Field/FiniteVolumeField.tpp:129-182,208-227  Field/VectorFiniteVolumeField.cpp:140-161
//   Field/ScalarGradient.cpp:34-74         Equation/FiniteVolumeEquation.tpp:64-86
#include <algorithm>
#include <cmath>

#include "comm.cuh"
#include "fv.cuh"
#include "kernels.cuh"
#include "solver.cuh"

using namespace phb;

namespace {

constexpr int kThreads = 256;

struct MeshView {
  SellView A;
  const int *linkFace;
  const double *vol, *fSx, *fSy, *fG, *fW, *fQx, *fQy;
  const int *fL, *fR;
  int nDev, nFaces;
  int nBCells;
  const int *bcCell, *bcPtr, *bcFace;
};

MeshView view(const phb_mesh *m) {
  MeshView v;
  v.A.sliceOff = m->sell.sliceOff.p; v.A.col = m->sell.col.p;
  v.A.nRows = m->sell.nRows; v.A.nSlices = m->sell.nSlices; v.A.nCols = m->sell.nCols;
  v.linkFace = m->dLinkFace.p;
  v.vol = m->dVol.p; v.fSx = m->dFSx.p; v.fSy = m->dFSy.p; v.fG = m->dFG.p; v.fW = m->dFW.p;
  v.fQx = m->dFQx.p; v.fQy = m->dFQy.p; v.fL = m->dFL.p; v.fR = m->dFR.p;
  v.nDev = m->nDev; v.nFaces = m->nFaces; v.nBCells = m->nBCells;
  v.bcCell = m->dBcCell.p; v.bcPtr = m->dBcPtr.p; v.bcFace = m->dBcFace.p;
  return v;
}

int row_grid(const phb_ctx *c, const phb_mesh *m) {
  const long long blocks = ((long long)m->sell.nSlices * 32 + kThreads - 1) / kThreads;
  return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)c->numSMs * 8));
}
int flat_grid(const phb_ctx *c, long long n) {
  const long long blocks = (n + kThreads - 1) / kThreads;
  return (int)std::max<long long>(1, std::min<long long>(blocks, (long long)c->numSMs * 8));
}

// row iteration shared by all cell-parallel kernels: warp <-> slice, lane <-> row
#define FOR_EACH_ROW(M)                                                                   \
  const int lane__ = threadIdx.x & 31;                                                    \
  const int wpb__ = blockDim.x >> 5;                                                      \
  for (int slice = blockIdx.x * wpb__ + (threadIdx.x >> 5); slice < (M).A.nSlices;        \
       slice += gridDim.x * wpb__) {                                                      \
    const int off = (M).A.sliceOff[slice];                                                \
    const int wdt = ((M).A.sliceOff[slice + 1] - off) >> 5;                               \
    const int row = slice * 32 + lane__;                                                  \
    const size_t slot0 = (size_t)off + lane__;                                            \
    if (row < (M).A.nRows) {
#define END_FOR_EACH_ROW }}

// ------------------------------------------------------------------ field glue
template <int NC>
__global__ void k_interp_faces(int nIF, const int *__restrict__ ifFace, MeshView M, int ldc, int ldf,
                               const double *__restrict__ cells, double *__restrict__ faces) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nIF; i += gridDim.x * blockDim.x) {
    const int f = ifFace[i];
    const double g = M.fW[f];
    const int l = M.fL[f], r = M.fR[f];
#pragma unroll
    for (int c = 0; c < NC; ++c)
      faces[(size_t)c * ldf + f] = g * cells[(size_t)c * ldc + l] + (1. - g) * cells[(size_t)c * ldc + r];
  }
}

template <int NC>
__global__ void k_boundary_faces(int nBF, const int *__restrict__ bfFace, const int *__restrict__ bfCell,
                                 const int *__restrict__ bfType, MeshView M, int ldc, int ldf,
                                 const double *__restrict__ cells, double *__restrict__ faces) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nBF) return;
  const int f = bfFace[i], l = bfCell[i], t = bfType[i];
  if (t == PHB_FIXED) return;
  if (NC == 2 && t == PHB_SYMMETRY) {
    const double nx = M.fSx[f], ny = M.fSy[f];
    const double ux = cells[l], uy = cells[(size_t)ldc + l];
    const double d = ux * nx + uy * ny, mm = nx * nx + ny * ny;
    faces[f] = ux - d * nx / mm;
    faces[(size_t)ldf + f] = uy - d * ny / mm;
    return;
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) faces[(size_t)c * ldf + f] = cells[(size_t)c * ldc + l];
}

// face gradient of a scalar: (phi_r - phi_l) r/|r|^2, boundary (phi_f - phi_l) r_f/|r_f|^2
__global__ void k_grad_faces(MeshView M, const double *__restrict__ phiC, const double *__restrict__ phiF,
                             double *__restrict__ gF) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < M.nFaces; f += gridDim.x * blockDim.x) {
    const int l = M.fL[f], r = M.fR[f];
    const double d = (r >= 0 ? phiC[r] : phiF[f]) - phiC[l];
    gF[f] = d * M.fQx[f];
    gF[(size_t)M.nFaces + f] = d * M.fQy[f];
  }
}

__device__ __forceinline__ void grad_links(const MeshView &M, size_t slot0, int wdt, const double *gF,
                                           double &tx, double &ty, double &sx, double &sy) {
  for (int k = 1; k < wdt; ++k) {
    const int lf = M.linkFace[slot0 + (size_t)k * 32];
    if (lf < 0) continue;
    const int f = lf >> 1;
    const double ax = fabs(M.fSx[f]), ay = fabs(M.fSy[f]);
    tx += gF[f] * ax;
    ty += gF[(size_t)M.nFaces + f] * ay;
    sx += ax;
    sy += ay;
  }
}
// cell gradient, FACE_TO_CELL: component-wise sum(g_f |S_f|) / sum |S_f|
__global__ void k_grad_cells(MeshView M, const double *__restrict__ gF, double *__restrict__ gC) {
  FOR_EACH_ROW(M)
    double tx = 0., ty = 0., sx = 0., sy = 0.;
    grad_links(M, slot0, wdt, gF, tx, ty, sx, sy);
    gC[row] = tx / sx;
    gC[(size_t)M.nDev + row] = ty / sy;
  END_FOR_EACH_ROW
}
__global__ void k_grad_cells_bnd(MeshView M, const double *__restrict__ gF, double *__restrict__ gC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  const int off = M.A.sliceOff[row >> 5];
  const int wdt = (M.A.sliceOff[(row >> 5) + 1] - off) >> 5;
  double tx = 0., ty = 0., sx = 0., sy = 0.;
  grad_links(M, (size_t)off + (row & 31), wdt, gF, tx, ty, sx, sy);
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    const double ax = fabs(M.fSx[f]), ay = fabs(M.fSy[f]);
    tx += gF[f] * ax;
    ty += gF[(size_t)M.nFaces + f] * ay;
    sx += ax;
    sy += ay;
  }
  gC[row] = tx / sx;
  gC[(size_t)M.nDev + row] = ty / sy;
}

// y[c][i] += a * x[c][i]
__global__ void k_axpy(long long n, int nc, long long ldy, long long ldx, double a,
                       const double *__restrict__ x, double *__restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * nc;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / n, j = i - c * n;
    y[c * ldy + j] += a * x[c * ldx + j];
  }
}
__global__ void k_fill(long long n, double v, double *__restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = v;
}
__global__ void k_fill_bfaces(int nBF, const int *__restrict__ bfFace, const int *__restrict__ bfPatch,
                              int nPatch, const double *__restrict__ ref, int nc, int ldf,
                              double *__restrict__ faces) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nBF) return;
  const int p = bfPatch[i];
  if (p < 0 || p >= nPatch) return;
  for (int c = 0; c < nc; ++c) faces[(size_t)c * ldf + bfFace[i]] = ref[2 * p + c];
}
__global__ void k_pack_field(int nSend, int nc, int ld, const int *__restrict__ sendDev,
                             const double *__restrict__ x, double *__restrict__ buf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nSend * nc) return;
  const int c = i / nSend, j = i - c * nSend;
  buf[i] = x[(size_t)c * ld + sendDev[j]];
}

// ------------------------------------------------------------------ assembly
// ddt: A_PP += rho V/dt ; rhs_P += -rho0 V phi0_P/dt
template <int NC>
__global__ void k_ddt(MeshView M, double *__restrict__ vals, double *__restrict__ rhs, int ldr,
                      const double *__restrict__ phi0, int ldc, double rhoConst,
                      const double *__restrict__ rho, const double *__restrict__ rho0, double dt, double sign,
                      const int *__restrict__ mask) {
  FOR_EACH_ROW(M)
    if (mask && !mask[row]) continue;
    const double V = M.vol[row];
    const double rh = rho ? rho[row] : rhoConst, rh0 = rho0 ? rho0[row] : rhoConst;
    vals[slot0] += sign * (rh * V / dt);
#pragma unroll
    for (int c = 0; c < NC; ++c)
      rhs[(size_t)c * ldr + row] += sign * (-rh0 * V * phi0[(size_t)c * ldc + row] / dt);
  END_FOR_EACH_ROW
}

// div (upwind, theta) / dive (MODE 1: explicit two-level form)
template <int NC, int MODE>
__global__ void k_div(MeshView M, double *__restrict__ vals, double *__restrict__ rhs, int ldr,
                      const double *__restrict__ uF, const double *__restrict__ u0F, const double *__restrict__ u1F,
                      const double *__restrict__ phi0, const double *__restrict__ phi1, int ldc,
                      double theta, double sign) {
  FOR_EACH_ROW(M)
    double diag = 0., r[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) r[c] = 0.;
    for (int k = 1; k < wdt; ++k) {
      const size_t slot = slot0 + (size_t)k * 32;
      const int lf = M.linkFace[slot];
      if (lf < 0) continue;
      const int f = lf >> 1, nb = M.A.col[slot];
      const double sg = (lf & 1) ? -1. : 1.;
      const double sx = M.fSx[f], sy = M.fSy[f];
      const double flux0 = sg * (u0F[f] * sx + u0F[(size_t)M.nFaces + f] * sy);
      if (MODE == 0) {
        if (theta != 0.) {
          const double flux = sg * (uF[f] * sx + uF[(size_t)M.nFaces + f] * sy);
          diag += theta * fmax(flux, 0.);
          vals[slot] += sign * (theta * fmin(flux, 0.));
        }
        const double a = (1. - theta) * fmax(flux0, 0.), b = (1. - theta) * fmin(flux0, 0.);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          r[c] += a * phi0[(size_t)c * ldc + row];
          r[c] += b * phi0[(size_t)c * ldc + nb];
        }
      } else {
        const double flux1 = sg * (u1F[f] * sx + u1F[(size_t)M.nFaces + f] * sy);
        const double a0 = theta * fmax(flux0, 0.), b0 = theta * fmin(flux0, 0.);
        const double a1 = (1. - theta) * fmax(flux1, 0.), b1 = (1. - theta) * fmin(flux1, 0.);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          r[c] += a0 * phi0[(size_t)c * ldc + row];
          r[c] += b0 * phi0[(size_t)c * ldc + nb];
          r[c] += a1 * phi1[(size_t)c * ldc + row];
          r[c] += b1 * phi1[(size_t)c * ldc + nb];
        }
      }
    }
    if (MODE == 0 && theta != 0.) vals[slot0] += sign * diag;
#pragma unroll
    for (int c = 0; c < NC; ++c) rhs[(size_t)c * ldr + row] += sign * r[c];
  END_FOR_EACH_ROW
}
template <int NC, int MODE>
__global__ void k_div_bnd(MeshView M, const int *__restrict__ faceType, double *__restrict__ vals,
                          double *__restrict__ rhs, int ldr, const double *__restrict__ uF,
                          const double *__restrict__ u0F, const double *__restrict__ u1F,
                          const double *__restrict__ phiF, const double *__restrict__ phi0F,
                          const double *__restrict__ phi1F, const double *__restrict__ phi0, int ldc,
                          double theta, double sign) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  const size_t slot0 = (size_t)M.A.sliceOff[row >> 5] + (row & 31);
  const size_t F = M.nFaces;
  double diag = 0., r[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) r[c] = 0.;
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    const int t = faceType[f];
    const double sx = M.fSx[f], sy = M.fSy[f];
    const double flux0 = u0F[f] * sx + u0F[F + f] * sy;
    if (MODE == 0) {
      const double flux = uF[f] * sx + uF[F + f] * sy;
      if (t == PHB_FIXED) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          r[c] += theta * flux * phiF[c * F + f];
          r[c] += (1. - theta) * flux0 * phi0F[c * F + f];
        }
      } else if (t == PHB_NORMAL_GRADIENT) {
        diag += theta * flux;
#pragma unroll
        for (int c = 0; c < NC; ++c) r[c] += (1. - theta) * flux0 * phi0[(size_t)c * ldc + row];
      }
    } else {
      const double flux1 = u1F[f] * sx + u1F[F + f] * sy;
      if (t == PHB_FIXED || t == PHB_NORMAL_GRADIENT) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          r[c] += theta * flux0 * phi0F[c * F + f];
          r[c] += (1. - theta) * flux1 * phi1F[c * F + f];
        }
      }
    }
  }
  if (MODE == 0) vals[slot0] += sign * diag;
#pragma unroll
  for (int c = 0; c < NC; ++c) rhs[(size_t)c * ldr + row] += sign * r[c];
}

// laplacian: c = Gamma_f g_f ; steady (theta < 0) or theta-weighted with old-time terms
template <int NC>
__global__ void k_lap(MeshView M, double *__restrict__ vals, double *__restrict__ rhs, int ldr,
                      double gammaConst, const double *__restrict__ gamF, const double *__restrict__ gam0F,
                      const double *__restrict__ phi0, int ldc, double theta, double sign) {
  const bool steady = theta < 0.;
  const double th = steady ? 1. : theta;
  FOR_EACH_ROW(M)
    double diag = 0., r[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) r[c] = 0.;
    for (int k = 1; k < wdt; ++k) {
      const size_t slot = slot0 + (size_t)k * 32;
      const int lf = M.linkFace[slot];
      if (lf < 0) continue;
      const int f = lf >> 1;
      const double g = M.fG[f];
      const double coeff = (gamF ? gamF[f] : gammaConst) * g;
      vals[slot] += sign * (th * coeff);
      diag -= th * coeff;
      if (!steady) {
        const int nb = M.A.col[slot];
        const double a = (1. - theta) * ((gam0F ? gam0F[f] : gammaConst) * g);
#pragma unroll
        for (int c = 0; c < NC; ++c) r[c] += a * (phi0[(size_t)c * ldc + nb] - phi0[(size_t)c * ldc + row]);
      }
    }
    vals[slot0] += sign * diag;
    if (!steady) {
#pragma unroll
      for (int c = 0; c < NC; ++c) rhs[(size_t)c * ldr + row] += sign * r[c];
    }
  END_FOR_EACH_ROW
}
template <int NC>
__global__ void k_lap_bnd(MeshView M, const int *__restrict__ faceType, double *__restrict__ vals,
                          double *__restrict__ rhs, int ldr, double gammaConst, const double *__restrict__ gamF,
                          const double *__restrict__ gam0F, const double *__restrict__ phiF,
                          const double *__restrict__ phi0F, const double *__restrict__ phi0, int ldc,
                          double theta, double sign, double *__restrict__ tens) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const bool steady = theta < 0.;
  const double th = steady ? 1. : theta;
  const int row = M.bcCell[i];
  const size_t slot0 = (size_t)M.A.sliceOff[row >> 5] + (row & 31);
  const size_t F = M.nFaces;
  double diag = 0., r[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) r[c] = 0.;
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    if (NC == 2 && !steady && tens && faceType[f] == PHB_SYMMETRY) {
      // tangential slip (UD/Laplacian.cpp:33-41, 95-101): -theta c on the shared diagonal, + theta c tw (x) tw on the
      // (cell, cell) block, source (1 - theta) c0 ((phi0 . tw) tw - phi0); tw = unit tangent of the face
      const double g = M.fG[f];
      const double coeff = (gamF ? gamF[f] : gammaConst) * g, coeff0 = (gam0F ? gam0F[f] : gammaConst) * g;
      const double sm = sqrt(M.fSx[f] * M.fSx[f] + M.fSy[f] * M.fSy[f]);
      const double tx = -M.fSy[f] / sm, ty = M.fSx[f] / sm;
      diag -= th * coeff;
      const size_t nL = (size_t)ldr;
      tens[row] += sign * th * coeff * tx * tx;
      tens[nL + row] += sign * th * coeff * tx * ty;
      tens[2 * nL + row] += sign * th * coeff * ty * tx;
      tens[3 * nL + row] += sign * th * coeff * ty * ty;
      const double p0x = phi0[row], p0y = phi0[(size_t)ldc + row], d = p0x * tx + p0y * ty;
      r[0] += (1. - theta) * coeff0 * (d * tx - p0x);
      r[NC - 1] += (1. - theta) * coeff0 * (d * ty - p0y);
      continue;
    }
    if (faceType[f] != PHB_FIXED) continue;
    const double g = M.fG[f];
    const double coeff = (gamF ? gamF[f] : gammaConst) * g;
    diag -= th * coeff;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      r[c] += th * coeff * phiF[c * F + f];
      if (!steady)
        r[c] += (1. - theta) * ((gam0F ? gam0F[f] : gammaConst) * g) *
                (phi0F[c * F + f] - phi0[(size_t)c * ldc + row]);
    }
  }
  vals[slot0] += sign * diag;
#pragma unroll
  for (int c = 0; c < NC; ++c) rhs[(size_t)c * ldr + row] += sign * r[c];
}

// The momentum predictor of FractionalStep in ONE pass over the rows (US/FractionalStep.cpp:82-83):
//   fv::ddt(u, dt) + fv::div(u, u, 0) == fv::laplacian(gamma, u, 0.5) - src::src(gradP)
// i.e. k_ddt + k_div<NC,0>(theta 0) + k_lap(theta 0.5, sign -1) + k_src on a zeroed equation, term sums kept apart
// and added in that order.  Every slot and rhs entry of a live row is WRITTEN (no zero-fill, no read-modify-write):
// 4 launches and 3 extra sweeps over `vals` less.  Boundary links follow in k_div_bnd / k_lap_bnd as before.
constexpr int kFusedBlocksPerSM = 6;   // 40 registers without spills; the grid is sized to be resident in one wave
template <int NC>
__global__ void __launch_bounds__(kThreads, kFusedBlocksPerSM) k_momentum_fused(MeshView M, double *__restrict__ vals, double *__restrict__ rhs, int ldr,
                                 const double *__restrict__ u0F, const double *__restrict__ phi0, int ldc,
                                 const double *__restrict__ gradP, double gamma, double dt) {
  FOR_EACH_ROW(M)
    const double V = M.vol[row];
    double p0[NC], rDiv[NC], rLap[NC], diagLap = 0.;
#pragma unroll
    for (int c = 0; c < NC; ++c) { p0[c] = phi0[(size_t)c * ldc + row]; rDiv[c] = 0.; rLap[c] = 0.; }
    for (int k = 1; k < wdt; ++k) {
      const size_t slot = slot0 + (size_t)k * 32;
      const int lf = M.linkFace[slot];
      if (lf < 0) { vals[slot] = 0.; continue; }
      const int f = lf >> 1, nb = M.A.col[slot];
      const double sg = (lf & 1) ? -1. : 1.;
      const double flux0 = sg * (u0F[f] * M.fSx[f] + u0F[(size_t)M.nFaces + f] * M.fSy[f]);
      const double a = fmax(flux0, 0.), b = fmin(flux0, 0.);
      const double coeff = gamma * M.fG[f];
      vals[slot] = -(0.5 * coeff);
      diagLap -= 0.5 * coeff;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double pn = phi0[(size_t)c * ldc + nb];
        rDiv[c] += a * p0[c];
        rDiv[c] += b * pn;
        rLap[c] += (0.5 * coeff) * (pn - p0[c]);
      }
    }
    vals[slot0] = V / dt - diagLap;
#pragma unroll
    for (int c = 0; c < NC; ++c)
      rhs[(size_t)c * ldr + row] = ((-V * p0[c] / dt + rDiv[c]) - rLap[c]) + gradP[(size_t)c * ldc + row] * V;
  END_FOR_EACH_ROW
}

// pEqn_ of FractionalStep in one pass (US/FractionalStep.cpp:97):  fv::laplacian(dt, p) == src::div(u), i.e. the steady
// k_lap<1> (sign +1) + k_flux_sum<0> (sign -1) on a zeroed equation; every slot and rhs entry of a live row is written.
__global__ void k_pressure_fused(MeshView M, double *__restrict__ vals, double *__restrict__ rhs,
                                 const double *__restrict__ uF, double gamma) {
  FOR_EACH_ROW(M)
    double diag = 0., d = 0.;
    for (int k = 1; k < wdt; ++k) {
      const size_t slot = slot0 + (size_t)k * 32;
      const int lf = M.linkFace[slot];
      if (lf < 0) { vals[slot] = 0.; continue; }
      const int f = lf >> 1;
      const double coeff = gamma * M.fG[f];
      vals[slot] = coeff;
      diag -= coeff;
      d += ((lf & 1) ? -1. : 1.) * (uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f]);
    }
    vals[slot0] = diag;
    rhs[row] = -d;
  END_FOR_EACH_ROW
}

// src::laplacian: rhs[row] += sign * sum c (phi_nb - phi_P)
__global__ void k_src_lap(MeshView M, double *__restrict__ rhs, double gammaConst, const double *__restrict__ gamF,
                          const double *__restrict__ phi, double sign) {
  FOR_EACH_ROW(M)
    double t = 0.;
    const double p0 = phi[row];
    for (int k = 1; k < wdt; ++k) {
      const size_t slot = slot0 + (size_t)k * 32;
      const int lf = M.linkFace[slot];
      if (lf < 0) continue;
      const int f = lf >> 1;
      t += (phi[M.A.col[slot]] - p0) * ((gamF ? gamF[f] : gammaConst) * M.fG[f]);
    }
    rhs[row] += sign * t;
  END_FOR_EACH_ROW
}
__global__ void k_src_lap_bnd(MeshView M, double *__restrict__ rhs, double gammaConst, const double *__restrict__ gamF,
                              const double *__restrict__ phi, const double *__restrict__ phiF, double sign) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  double t = 0.;
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    t += (phiF[f] - phi[row]) * ((gamF ? gamF[f] : gammaConst) * M.fG[f]);
  }
  rhs[row] += sign * t;
}

// src::src: rhs += sign * f V
template <int NC>
__global__ void k_src(MeshView M, double *__restrict__ rhs, int ldr, const double *__restrict__ f, int ldc,
                      double sign) {
  FOR_EACH_ROW(M)
    (void)wdt; (void)slot0;
#pragma unroll
    for (int c = 0; c < NC; ++c) rhs[(size_t)c * ldr + row] += sign * (f[(size_t)c * ldc + row] * M.vol[row]);
  END_FOR_EACH_ROW
}

// sum_f u_f . S_out over the links of a cell.  OUT 0: rhs += sign*div ; OUT 1: out = div ;
// OUT 2: out = dt/V * sum max(flux,0) (Courant number)
template <int OUT>
__global__ void k_flux_sum(MeshView M, const double *__restrict__ uF, double *__restrict__ out, double sign,
                           double dt, const int *__restrict__ mask) {
  FOR_EACH_ROW(M)
    if (mask && !mask[row]) continue;
    double d = 0.;
    for (int k = 1; k < wdt; ++k) {
      const int lf = M.linkFace[slot0 + (size_t)k * 32];
      if (lf < 0) continue;
      const int f = lf >> 1;
      const double flux = ((lf & 1) ? -1. : 1.) * (uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f]);
      d += (OUT == 2) ? fmax(flux, 0.) : flux;
    }
    if (OUT == 0) out[row] += sign * d;
    else out[row] = d;
  END_FOR_EACH_ROW
}
template <int OUT>
__global__ void k_flux_sum_bnd(MeshView M, const double *__restrict__ uF, double *__restrict__ out, double sign,
                               double dt, const int *__restrict__ mask) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  if (mask && !mask[row]) return;
  double d = 0.;
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    const double flux = uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f];
    d += (OUT == 2) ? fmax(flux, 0.) : flux;
  }
  if (OUT == 0) out[row] += sign * d;
  else out[row] += d;
}
// both per-step diagnostics of FractionalStep::solve (US/FractionalStep.cpp:41-43) in one pass over the links:
// div[row] = sum_f u_f . S_out, co[row] = sum_f max(u_f . S_out, 0)
__global__ void k_flux_both(MeshView M, const double *__restrict__ uF, double *__restrict__ div, double *__restrict__ co) {
  FOR_EACH_ROW(M)
    double d = 0., p = 0.;
    for (int k = 1; k < wdt; ++k) {
      const int lf = M.linkFace[slot0 + (size_t)k * 32];
      if (lf < 0) continue;
      const int f = lf >> 1;
      const double flux = ((lf & 1) ? -1. : 1.) * (uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f]);
      d += flux;
      p += fmax(flux, 0.);
    }
    div[row] = d;
    co[row] = p;
  END_FOR_EACH_ROW
}
__global__ void k_flux_both_bnd(MeshView M, const double *__restrict__ uF, double *__restrict__ div, double *__restrict__ co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  double d = 0., p = 0.;
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    const double flux = uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f];
    d += flux;
    p += fmax(flux, 0.);
  }
  div[row] += d;
  co[row] += p;
}
// out[0] = max |div|, out[1] = max co dt / V over owned rows
__global__ void k_max_reduce2(int n, const double *__restrict__ div, const double *__restrict__ co,
                              const double *__restrict__ vol, double dt, double *partials, unsigned *ticket, double *out) {
  double m0 = 0., m1 = 0.;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    m0 = fmax(m0, fabs(div[i]));
    m1 = fmax(m1, co[i] * dt / vol[i]);
  }
  double v[2] = {m0, m1};
  grid_reduce<2, true>(v, partials, ticket, out);
}

// max |x| (MODE 0) or max (x dt / V) (MODE 1) over owned rows
template <int MODE>
__global__ void k_max_reduce(int n, const double *__restrict__ x, const double *__restrict__ vol, double dt,
                             double *partials, unsigned *ticket, double *out) {
  double mx = 0.;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    mx = fmax(mx, MODE == 0 ? fabs(x[i]) : x[i] * dt / vol[i]);
  double v[1] = {mx};
  grid_reduce<1, true>(v, partials, ticket, out);
}

// rho * eqn: scale row coefficients and rhs (UE/VectorFiniteVolumeEquation.cpp:163-170)
template <int NC>
__global__ void k_scale_rows(MeshView M, double *__restrict__ vals, double *__restrict__ rhs, int ldr,
                             const double *__restrict__ rho) {
  FOR_EACH_ROW(M)
    const double s = rho[row];
    for (int k = 0; k < wdt; ++k) vals[slot0 + (size_t)k * 32] *= s;
#pragma unroll
    for (int c = 0; c < NC; ++c) rhs[(size_t)c * ldr + row] *= s;
  END_FOR_EACH_ROW
}
__global__ void k_scale_tens(int n, const double *__restrict__ rho, double *__restrict__ tens) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * n; i += gridDim.x * blockDim.x) tens[i] *= rho[i % n];
}
// relax(omega): a_PP /= omega ; rhs_P -= (1-omega) a_PP phi_P
template <int NC>
__global__ void k_relax(MeshView M, double *__restrict__ vals, double *__restrict__ rhs, int ldr,
                        const double *__restrict__ phi, int ldc, double omega) {
  FOR_EACH_ROW(M)
    (void)wdt;
    const double a = vals[slot0] / omega;
    vals[slot0] = a;
#pragma unroll
    for (int c = 0; c < NC; ++c) rhs[(size_t)c * ldr + row] -= (1. - omega) * a * phi[(size_t)c * ldc + row];
  END_FOR_EACH_ROW
}
// b = -rhs into the solver's vector (leading dimension ld)
__global__ void k_neg_copy(int n, int nc, int ldr, int ld, const double *__restrict__ rhs,
                           double *__restrict__ b) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n * nc;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / n, j = i - c * n;
    b[c * ld + j] = -rhs[c * ldr + j];
  }
}
__global__ void k_copy2d(int n, int nc, long long ldSrc, long long ldDst, const double *__restrict__ src,
                         double *__restrict__ dst) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < (long long)n * nc;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / n, j = i - c * n;
    dst[c * ldDst + j] = src[c * ldSrc + j];
  }
}

int check_pair(const phb_eqn *e, const phb_field *f, const char *what) {
  PHB_REQUIRE(e && f, "%s: NULL argument", what);
  PHB_REQUIRE(e->m == f->m, "%s: field and equation live on different meshes", what);
  return PHB_OK;
}

}  // namespace

namespace phb {

// per-FACE boundary type of a field (0 for interior faces), rebuilt when BCs change
int field_face_types(phb_field *f) {
  if (!f->bcDirty) return PHB_OK;
  phb_mesh *m = f->m;
  std::vector<int> ft(m->nFaces, PHB_NORMAL_GRADIENT), bt;
  std::vector<double> ref(2 * std::max<size_t>(1, f->bc.size()), 0.);
  for (int fc = 0; fc < m->nFaces; ++fc) {
    const int p = m->fPatch[fc];
    if (m->fR[fc] < 0 && p >= 0 && p < (int)f->bc.size()) ft[fc] = f->bc[p].type;
    if (m->fR[fc] < 0) bt.push_back(ft[fc]);
  }
  PHB_CHECK(f->dFaceType.upload(ft, m->ctx->stream));
  PHB_CHECK(f->dBfType.upload(bt, m->ctx->stream));
  PHB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  f->bcDirty = false;
  return PHB_OK;
}

int field_interpolate_faces(phb_field *f) {
  phb_mesh *m = f->m;
  phb_ctx *c = m->ctx;
  const MeshView M = view(m);
  if (m->nIFaces) {
    if (f->nComp == 1)
      PHB_LAUNCH(c, k_interp_faces<1>, flat_grid(c, m->nIFaces), kThreads, 0, m->nIFaces, m->dIfFace.p, M,
                 m->nDev, m->nFaces, f->cells.p, f->faces.p);
    else
      PHB_LAUNCH(c, k_interp_faces<2>, flat_grid(c, m->nIFaces), kThreads, 0, m->nIFaces, m->dIfFace.p, M,
                 m->nDev, m->nFaces, f->cells.p, f->faces.p);
  }
  return field_set_boundary_faces(f);
}

int field_set_boundary_faces(phb_field *f) {
  phb_mesh *m = f->m;
  phb_ctx *c = m->ctx;
  PHB_CHECK(field_face_types(f));
  if (!m->nBFaces) return PHB_OK;
  const MeshView M = view(m);
  if (f->nComp == 1)
    PHB_LAUNCH(c, k_boundary_faces<1>, (m->nBFaces + 255) / 256, 256, 0, m->nBFaces, m->dBfFace.p, m->dBfCell.p,
               f->dBfType.p, M, m->nDev, m->nFaces, f->cells.p, f->faces.p);
  else
    PHB_LAUNCH(c, k_boundary_faces<2>, (m->nBFaces + 255) / 256, 256, 0, m->nBFaces, m->dBfFace.p, m->dBfCell.p,
               f->dBfType.p, M, m->nDev, m->nFaces, f->cells.p, f->faces.p);
  return PHB_OK;
}

int field_gradient(const phb_field *phi, phb_field *grad) {
  phb_mesh *m = phi->m;
  phb_ctx *c = m->ctx;
  const MeshView M = view(m);
  PHB_LAUNCH(c, k_grad_faces, flat_grid(c, m->nFaces), kThreads, 0, M, phi->cells.p, phi->faces.p, grad->faces.p);
  PHB_LAUNCH(c, k_grad_cells, row_grid(c, m), kThreads, 0, M, grad->faces.p, grad->cells.p);
  if (m->nBCells)
    PHB_LAUNCH(c, k_grad_cells_bnd, (m->nBCells + 255) / 256, 256, 0, M, grad->faces.p, grad->cells.p);
  return PHB_OK;
}

int field_axpy_cells(phb_field *y, double a, const phb_field *x) {
  phb_mesh *m = y->m;
  PHB_LAUNCH(m->ctx, k_axpy, flat_grid(m->ctx, (long long)m->nLocal * y->nComp), kThreads, 0,
             (long long)m->nLocal, y->nComp, (long long)m->nDev, (long long)m->nDev, a, x->cells.p, y->cells.p);
  return PHB_OK;
}
int field_axpy_faces(phb_field *y, double a, const phb_field *x) {
  phb_mesh *m = y->m;
  PHB_LAUNCH(m->ctx, k_axpy, flat_grid(m->ctx, (long long)m->nFaces * y->nComp), kThreads, 0,
             (long long)m->nFaces, y->nComp, (long long)m->nFaces, (long long)m->nFaces, a, x->faces.p, y->faces.p);
  return PHB_OK;
}

// true when no patch of the field is FIXED on ANY rank: the Laplacian of such a field is singular
int field_all_neumann(phb_field *f, bool *out) {
  phb_mesh *m = f->m;
  phb_ctx *c = m->ctx;
  double has = 0.;
  for (int fc = 0; fc < m->nFaces; ++fc) {
    const int p = m->fPatch[fc];
    if (m->fR[fc] < 0 && m->owner[m->fL[fc]] == m->rank && p >= 0 && p < (int)f->bc.size() && f->bc[p].type == PHB_FIXED) {
      has = 1.;
      break;
    }
  }
  if (c->nProcs > 1) {
    phb::DevBuf<double> d;
    PHB_CHECK(d.upload(&has, 1, c->stream));
    PHB_CHECK(comm_allreduce_max(c, d.p, 1));
    PHB_CUDA(cudaMemcpyAsync(&has, d.p, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    PHB_CUDA(cudaStreamSynchronize(c->stream));
  }
  *out = has == 0.;
  return PHB_OK;
}

int field_send_messages(phb_field *f) {
  phb_mesh *m = f->m;
  phb_ctx *c = m->ctx;
  if (c->nProcs == 1) return PHB_OK;
  const int nSend = (int)m->hSendDev.size();
  if (nSend)
    PHB_LAUNCH(c, k_pack_field, (nSend * f->nComp + 255) / 256, 256, 0, nSend, f->nComp, m->nDev, m->dSendDev.p,
               f->cells.p, m->dSendBuf.p);
  for (int k = 0; k < f->nComp; ++k)
    PHB_CHECK(comm_exchange(c, m->dSendBuf.p + (size_t)k * nSend, m->hSendOff.data(), m->hSendCnt.data(),
                            f->cells.p + (size_t)k * m->nDev, m->hRecvOff.data(), m->hRecvCnt.data()));
  return PHB_OK;
}

// out = max over owned cells; mode 0: |sum_f u_f.S_f|, mode 1: Courant number
int field_flux_max(const phb_field *u, int mode, double dt, phb::DevBuf<double> &scratch,
                   phb::DevBuf<double> &partials, phb::DevBuf<unsigned> &ticket, double *devOut) {
  phb_mesh *m = u->m;
  phb_ctx *c = m->ctx;
  const MeshView M = view(m);
  PHB_CHECK(scratch.alloc(2 * (size_t)m->nLocal));   // same sizes as field_flux_diagnostics: no reallocation between them
  const int grid = row_grid(c, m);
  if (mode == 0) {
    PHB_LAUNCH(c, k_flux_sum<1>, grid, kThreads, 0, M, u->faces.p, scratch.p, 1., dt, (const int *)nullptr);
    if (m->nBCells) PHB_LAUNCH(c, k_flux_sum_bnd<1>, (m->nBCells + 255) / 256, 256, 0, M, u->faces.p, scratch.p, 1., dt, (const int *)nullptr);
  } else {
    PHB_LAUNCH(c, k_flux_sum<2>, grid, kThreads, 0, M, u->faces.p, scratch.p, 1., dt, (const int *)nullptr);
    if (m->nBCells) PHB_LAUNCH(c, k_flux_sum_bnd<2>, (m->nBCells + 255) / 256, 256, 0, M, u->faces.p, scratch.p, 1., dt, (const int *)nullptr);
  }
  const int g2 = flat_grid(c, m->nLocal);
  PHB_CHECK(partials.alloc((size_t)c->numSMs * 8 * 2));
  if (!ticket.p) { PHB_CHECK(ticket.alloc(1)); PHB_CHECK(ticket.zero(c->stream)); }
  if (mode == 0)
    PHB_LAUNCH(c, k_max_reduce<0>, g2, kThreads, 0, m->nLocal, scratch.p, m->dVol.p, dt, partials.p, ticket.p, devOut);
  else
    PHB_LAUNCH(c, k_max_reduce<1>, g2, kThreads, 0, m->nLocal, scratch.p, m->dVol.p, dt, partials.p, ticket.p, devOut);
  PHB_CHECK(comm_allreduce_max(c, devOut, 1));
  return PHB_OK;
}

// devOut[0] = max |sum_f u_f . S_f| (divergence error), devOut[1] = max Courant number; one pass over the links
int field_flux_diagnostics(const phb_field *u, double dt, phb::DevBuf<double> &scratch, phb::DevBuf<double> &partials,
                           phb::DevBuf<unsigned> &ticket, double *devOut) {
  phb_mesh *m = u->m;
  phb_ctx *c = m->ctx;
  const MeshView M = view(m);
  const size_t n = (size_t)m->nLocal;
  PHB_CHECK(scratch.alloc(2 * n));
  double *div = scratch.p, *co = scratch.p + n;
  PHB_LAUNCH(c, k_flux_both, row_grid(c, m), kThreads, 0, M, u->faces.p, div, co);
  if (m->nBCells) PHB_LAUNCH(c, k_flux_both_bnd, (m->nBCells + 255) / 256, 256, 0, M, u->faces.p, div, co);
  PHB_CHECK(partials.alloc((size_t)c->numSMs * 8 * 2));
  if (!ticket.p) { PHB_CHECK(ticket.alloc(1)); PHB_CHECK(ticket.zero(c->stream)); }
  PHB_LAUNCH(c, k_max_reduce2, flat_grid(c, m->nLocal), kThreads, 0, m->nLocal, (const double *)div, (const double *)co,
             m->dVol.p, dt, partials.p, ticket.p, devOut);
  PHB_CHECK(comm_allreduce_max(c, devOut, 2));
  return PHB_OK;
}

}  // namespace phb

// ------------------------------------------------------------------- C ABI
extern "C" {

int phb_field_create(phb_mesh *m, int nComp, const char *name, phb_field **out) {
  PHB_REQUIRE(m && out && (nComp == 1 || nComp == 2), "phb_field_create: bad argument");
  PHB_REQUIRE(m->finalized, "phb_field_create: mesh is not finalized");
  if (m->ctx->device < 0) { phb::set_error("phb_field_create: host-only context"); return PHB_ERR_STATE; }
  std::unique_ptr<phb_field> f(new phb_field());
  f->m = m; f->nComp = nComp; f->name = name ? name : "";
  f->bc.assign(m->patchNames.size(), BcEntry());
  PHB_CHECK(f->cells.alloc((size_t)nComp * m->nDev)); PHB_CHECK(f->faces.alloc((size_t)nComp * m->nFaces));
  PHB_CHECK(f->cells0.alloc((size_t)nComp * m->nDev)); PHB_CHECK(f->faces0.alloc((size_t)nComp * m->nFaces));
  PHB_CHECK(f->cells.zero(m->ctx->stream)); PHB_CHECK(f->faces.zero(m->ctx->stream));
  PHB_CHECK(f->cells0.zero(m->ctx->stream)); PHB_CHECK(f->faces0.zero(m->ctx->stream));
  *out = f.release();
  return PHB_OK;
}
int phb_field_destroy(phb_field *f) { delete f; return PHB_OK; }

// setBoundaryTypes/RefValues: every face of the patch takes the reference value
int phb_field_set_bc(phb_field *f, const char *patch, int type, double vx, double vy) {
  PHB_REQUIRE(f && patch, "phb_field_set_bc: NULL argument");
  PHB_REQUIRE(type == PHB_FIXED || type == PHB_NORMAL_GRADIENT || type == PHB_SYMMETRY,
              "phb_field_set_bc: unrecognized boundary type %d", type);
  phb_mesh *m = f->m;
  f->bcVersion++;   // on every rank, whether or not the patch touches it: matrix tags must change everywhere at once
  const int p = phb_mesh_patch_id(m, patch);
  // a partitioned mesh only carries the patches that touch it: boundary input for
  // the others is ignored, as setBoundaryTypes does (UF/FiniteVolumeField.tpp:452-476)
  if (p < 0 && m->nProcs > 1) return PHB_OK;
  PHB_REQUIRE(p >= 0, "phb_field_set_bc: no patch named \"%s\"", patch);
  f->bc[p].type = type; f->bc[p].vx = vx; f->bc[p].vy = vy;
  f->bcDirty = true;
  std::vector<double> ref(2 * f->bc.size(), 0.);
  // only this patch's faces are touched
  std::vector<int> patchOnly(m->nBFaces, -1);
  int i = 0;
  for (int fc = 0; fc < m->nFaces; ++fc)
    if (m->fR[fc] < 0) { patchOnly[i] = m->fPatch[fc] == p ? p : -1; ++i; }
  ref[2 * p] = vx; ref[2 * p + 1] = vy;
  phb::DevBuf<int> dP;
  phb::DevBuf<double> dR;
  PHB_CHECK(dP.upload(patchOnly, m->ctx->stream));
  PHB_CHECK(dR.upload(ref, m->ctx->stream));
  if (m->nBFaces)
    PHB_LAUNCH(m->ctx, k_fill_bfaces, (m->nBFaces + 255) / 256, 256, 0, m->nBFaces, m->dBfFace.p, dP.p,
               (int)f->bc.size(), dR.p, f->nComp, m->nFaces, f->faces.p);
  PHB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  return PHB_OK;
}

// dir 0: dev[c][cell2dev[i]] = host[c][i]; dir 1: host[c][i] = dev[c][cell2dev[i]]
__global__ void k_cells_permute(int nCells, int nComp, int nDev, const int *__restrict__ cell2dev,
                                double *__restrict__ host, double *__restrict__ dev, int dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nCells) return;
  const int d = cell2dev[i];
  for (int c = 0; c < nComp; ++c) {
    if (dir == 0) dev[(size_t)c * nDev + d] = host[(size_t)c * nCells + i];
    else host[(size_t)c * nCells + i] = dev[(size_t)c * nDev + d];
  }
}

static int field_part(phb_field *f, const char *part, double **dev, long long *len, bool *isCells) {
  phb_mesh *m = f->m;
  const std::string p(part ? part : "");
  if (p == "cells") { *dev = f->cells.p; *isCells = true; }
  else if (p == "cells0") { *dev = f->cells0.p; *isCells = true; }
  else if (p == "faces") { *dev = f->faces.p; *isCells = false; }
  else if (p == "faces0") { *dev = f->faces0.p; *isCells = false; }
  else { phb::set_error("unknown field part \"%s\"", p.c_str()); return PHB_ERR_ARG; }
  *len = (long long)f->nComp * (*isCells ? m->nCells : m->nFaces);
  return PHB_OK;
}

int phb_field_set(phb_field *f, const char *part, const double *v, long long n) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(f && v, "phb_field_set: NULL argument");
  double *dev; long long len; bool isCells;
  PHB_CHECK(field_part(f, part, &dev, &len, &isCells));
  PHB_REQUIRE(n == len, "phb_field_set: size %lld != %lld", n, len);
  phb_mesh *m = f->m;
  if (isCells && !m->identityCells) {
    // host (reference) cell order -> device order, permuted on the device: no pageable staging copy
    PHB_CHECK(m->dStage.alloc((size_t)2 * m->nCells));
    PHB_CUDA(cudaMemcpyAsync(m->dStage.p, v, n * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
    PHB_LAUNCH(m->ctx, k_cells_permute, (m->nCells + 255) / 256, 256, 0, m->nCells, f->nComp, m->nDev,
               m->dCell2Dev.p, m->dStage.p, dev, 0);
    PHB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  } else {
    PHB_CUDA(cudaMemcpyAsync(dev, v, n * sizeof(double), cudaMemcpyHostToDevice, m->ctx->stream));
    PHB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  }
  return PHB_OK;
  PHB_TRY_END
}

int phb_field_get(const phb_field *cf, const char *part, double *v, long long n) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(cf && v, "phb_field_get: NULL argument");
  phb_field *f = const_cast<phb_field *>(cf);
  double *dev; long long len; bool isCells;
  PHB_CHECK(field_part(f, part, &dev, &len, &isCells));
  PHB_REQUIRE(n == len, "phb_field_get: size %lld != %lld", n, len);
  phb_mesh *m = f->m;
  if (isCells && !m->identityCells) {
    PHB_CHECK(m->dStage.alloc((size_t)2 * m->nCells));
    PHB_LAUNCH(m->ctx, k_cells_permute, (m->nCells + 255) / 256, 256, 0, m->nCells, f->nComp, m->nDev,
               m->dCell2Dev.p, m->dStage.p, dev, 1);
    PHB_CUDA(cudaMemcpyAsync(v, m->dStage.p, n * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    PHB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  } else {
    PHB_CUDA(cudaMemcpyAsync(v, dev, n * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
    PHB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  }
  return PHB_OK;
  PHB_TRY_END
}

int phb_field_fill(phb_field *f, double vx, double vy) {
  PHB_REQUIRE(f, "phb_field_fill: NULL argument");
  phb_mesh *m = f->m;
  const double v[2] = {vx, vy};
  for (int c = 0; c < f->nComp; ++c) {
    PHB_LAUNCH(m->ctx, k_fill, flat_grid(m->ctx, m->nDev), kThreads, 0, (long long)m->nDev, v[c],
               f->cells.p + (size_t)c * m->nDev);
    PHB_LAUNCH(m->ctx, k_fill, flat_grid(m->ctx, m->nFaces), kThreads, 0, (long long)m->nFaces, v[c],
               f->faces.p + (size_t)c * m->nFaces);
  }
  return PHB_OK;
}

// deep copy of cells + faces (the reference copies; a pointer swap would alias)
int phb_field_save_previous(phb_field *f) {
  PHB_REQUIRE(f, "phb_field_save_previous: NULL argument");
  cudaStream_t st = f->m->ctx->stream;
  PHB_CUDA(cudaMemcpyAsync(f->cells0.p, f->cells.p, f->cells.n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  PHB_CUDA(cudaMemcpyAsync(f->faces0.p, f->faces.p, f->faces.n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  f->hasOld = true;
  return PHB_OK;
}

int phb_field_interpolate_faces(phb_field *f) {
  PHB_REQUIRE(f, "phb_field_interpolate_faces: NULL argument");
  return phb::field_interpolate_faces(f);
}
int phb_field_set_boundary_faces(phb_field *f) {
  PHB_REQUIRE(f, "phb_field_set_boundary_faces: NULL argument");
  return phb::field_set_boundary_faces(f);
}
int phb_field_gradient(const phb_field *phi, phb_field *grad) {
  PHB_REQUIRE(phi && grad && phi->m == grad->m && phi->nComp == 1 && grad->nComp == 2,
              "phb_field_gradient: needs a scalar field and a vector field on the same mesh");
  return phb::field_gradient(phi, grad);
}
int phb_field_send_messages(phb_field *f) {
  PHB_REQUIRE(f, "phb_field_send_messages: NULL argument");
  return phb::field_send_messages(f);
}

int phb_eqn_create(phb_mesh *m, int nComp, phb_eqn **out) {
  PHB_REQUIRE(m && out && (nComp == 1 || nComp == 2), "phb_eqn_create: bad argument");
  PHB_REQUIRE(m->finalized, "phb_eqn_create: mesh is not finalized");
  if (m->ctx->device < 0) { phb::set_error("phb_eqn_create: host-only context"); return PHB_ERR_STATE; }
  std::unique_ptr<phb_eqn> e(new phb_eqn());
  e->m = m; e->nComp = nComp;
  PHB_CHECK(e->vals.alloc((size_t)m->sell.nSlots));
  PHB_CHECK(e->rhs.alloc((size_t)nComp * m->nLocal));
  PHB_CHECK(e->vals.zero(m->ctx->stream)); PHB_CHECK(e->rhs.zero(m->ctx->stream));
  *out = e.release();
  return PHB_OK;
}
int phb_eqn_destroy(phb_eqn *e) { delete e; return PHB_OK; }
int phb_eqn_zero(phb_eqn *e) {
  PHB_REQUIRE(e, "phb_eqn_zero: NULL argument");
  PHB_CHECK(e->vals.zero(e->m->ctx->stream));
  if (e->tens.p) PHB_CHECK(e->tens.zero(e->m->ctx->stream));
  e->hasTens = false;
  return e->rhs.zero(e->m->ctx->stream);
}

// 0 / 1 per owned row (device numbering) from a list of reference cell ids
static int cell_mask(phb_mesh *m, int nCells, const int *cells, phb::DevBuf<int> &mask) {
  std::vector<int> h(std::max(1, m->nLocal), 0);
  for (int i = 0; i < nCells; ++i) {
    PHB_REQUIRE(cells[i] >= 0 && cells[i] < m->nCells, "cell group: cell id %d out of range", cells[i]);
    const int d = m->cell2dev[cells[i]];
    if (d < m->nLocal) h[d] = 1;
  }
  PHB_CHECK(mask.upload(h, m->ctx->stream));
  return PHB_OK;
}

static int assemble_ddt(phb_eqn *e, const phb_field *phi, double rhoConst, const phb_field *rho, double dt,
                        double sign, const int *mask) {
  PHB_CHECK(check_pair(e, phi, "phb_assemble_ddt"));
  PHB_REQUIRE(phi->nComp == e->nComp, "phb_assemble_ddt: component mismatch");
  PHB_REQUIRE(phi->hasOld, "phb_assemble_ddt: field has no previous time step (savePreviousTimeStep)");
  PHB_REQUIRE(!rho || (rho->nComp == 1 && rho->m == e->m), "phb_assemble_ddt: rho must be a scalar field");
  PHB_REQUIRE(dt > 0., "phb_assemble_ddt: dt must be positive");
  phb_mesh *m = e->m;
  const MeshView M = view(m);
  const double *r = rho ? rho->cells.p : nullptr, *r0 = rho ? (rho->hasOld ? rho->cells0.p : rho->cells.p) : nullptr;
  if (e->nComp == 1)
    PHB_LAUNCH(m->ctx, k_ddt<1>, row_grid(m->ctx, m), kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, phi->cells0.p,
               m->nDev, rhoConst, r, r0, dt, sign, mask);
  else
    PHB_LAUNCH(m->ctx, k_ddt<2>, row_grid(m->ctx, m), kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, phi->cells0.p,
               m->nDev, rhoConst, r, r0, dt, sign, mask);
  return PHB_OK;
}
int phb_assemble_ddt(phb_eqn *e, const phb_field *phi, double rhoConst, const phb_field *rho, double dt,
                     double sign) {
  return assemble_ddt(e, phi, rhoConst, rho, dt, sign, nullptr);
}
// fv::ddt(field, timeStep, cells) (UD/TimeDerivative.h:50-62): the listed cells only (reference cell ids)
int phb_assemble_ddt_cells(phb_eqn *e, const phb_field *phi, double dt, double sign, int nCells, const int *cells) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(e && cells && nCells >= 0, "phb_assemble_ddt_cells: bad argument");
  phb::DevBuf<int> mask;
  PHB_CHECK(cell_mask(e->m, nCells, cells, mask));
  PHB_CHECK(assemble_ddt(e, phi, 1., nullptr, dt, sign, mask.p));
  PHB_CUDA(cudaStreamSynchronize(e->m->ctx->stream));   // the mask goes out of scope
  return PHB_OK;
  PHB_TRY_END
}

static int assemble_div(phb_eqn *e, const phb_field *u, const phb_field *cphi, double theta, double sign, int mode,
                        const char *what) {
  PHB_CHECK(check_pair(e, cphi, what));
  PHB_REQUIRE(u && u->m == e->m && u->nComp == 2, "%s: u must be a vector field on the same mesh", what);
  PHB_REQUIRE(cphi->nComp == e->nComp, "%s: component mismatch", what);
  PHB_REQUIRE(u->hasOld && cphi->hasOld, "%s: fields need a previous time step", what);
  phb_field *phi = const_cast<phb_field *>(cphi);
  PHB_CHECK(phb::field_face_types(phi));
  phb_mesh *m = e->m;
  phb_ctx *c = m->ctx;
  const MeshView M = view(m);
  const int grid = row_grid(c, m), gb = (m->nBCells + 255) / 256;
  // the reference keeps ONE old level aliased as oldField(0) and oldField(1)
  const double *u0F = u->faces0.p, *u1F = u->faces0.p;
  const double *p0 = phi->cells0.p, *p1 = phi->cells0.p, *p0F = phi->faces0.p, *p1F = phi->faces0.p;
#define DIV_LAUNCH(NC, MODE)                                                                              \
  do {                                                                                                    \
    PHB_LAUNCH(c, (k_div<NC, MODE>), grid, kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, u->faces.p,    \
               u0F, u1F, p0, p1, m->nDev, theta, sign);                                                   \
    if (m->nBCells)                                                                                       \
      PHB_LAUNCH(c, (k_div_bnd<NC, MODE>), gb, 256, 0, M, phi->dFaceType.p, e->vals.p, e->rhs.p,          \
                 m->nLocal, u->faces.p, u0F, u1F, phi->faces.p, p0F, p1F, p0, m->nDev, theta, sign);      \
  } while (0)
  if (e->nComp == 1) { if (mode == 0) DIV_LAUNCH(1, 0); else DIV_LAUNCH(1, 1); }
  else { if (mode == 0) DIV_LAUNCH(2, 0); else DIV_LAUNCH(2, 1); }
#undef DIV_LAUNCH
  return PHB_OK;
}
int phb_assemble_div(phb_eqn *e, const phb_field *u, const phb_field *phi, double theta, double sign) {
  return assemble_div(e, u, phi, theta, sign, 0, "phb_assemble_div");
}
int phb_assemble_dive(phb_eqn *e, const phb_field *u, const phb_field *phi, double theta, double sign) {
  return assemble_div(e, u, phi, theta, sign, 1, "phb_assemble_dive");
}

int phb_assemble_laplacian(phb_eqn *e, double gammaConst, const phb_field *gam, const phb_field *cphi,
                           double theta, double sign) {
  PHB_CHECK(check_pair(e, cphi, "phb_assemble_laplacian"));
  PHB_REQUIRE(cphi->nComp == e->nComp, "phb_assemble_laplacian: component mismatch");
  PHB_REQUIRE(!gam || (gam->nComp == 1 && gam->m == e->m), "phb_assemble_laplacian: gamma must be a scalar field");
  PHB_REQUIRE(theta < 0. || cphi->hasOld, "phb_assemble_laplacian: theta form needs a previous time step");
  phb_field *phi = const_cast<phb_field *>(cphi);
  phb_mesh *m = e->m;
  phb_ctx *c = m->ctx;
  double *tens = nullptr;
  if (phi->nComp == 2 && theta >= 0.)     // the steady overloads are the primary templates: SYMMETRY adds nothing (UD/Laplacian.h:35-37)
    for (const BcEntry &b : phi->bc)
      if (b.type == PHB_SYMMETRY) {
        if (!e->tens.p) { PHB_CHECK(e->tens.alloc(4 * (size_t)m->nLocal)); PHB_CHECK(e->tens.zero(c->stream)); }
        e->hasTens = true;
        tens = e->tens.p;
      }
  PHB_CHECK(phb::field_face_types(phi));
  const MeshView M = view(m);
  const double *gF = gam ? gam->faces.p : nullptr;
  const double *g0F = gam ? (gam->hasOld ? gam->faces0.p : gam->faces.p) : nullptr;
  const int grid = row_grid(c, m), gb = (m->nBCells + 255) / 256;
  if (e->nComp == 1) {
    PHB_LAUNCH(c, k_lap<1>, grid, kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, gammaConst, gF, g0F, phi->cells0.p,
               m->nDev, theta, sign);
    if (m->nBCells)
      PHB_LAUNCH(c, k_lap_bnd<1>, gb, 256, 0, M, phi->dFaceType.p, e->vals.p, e->rhs.p, m->nLocal, gammaConst, gF,
                 g0F, phi->faces.p, phi->faces0.p, phi->cells0.p, m->nDev, theta, sign, (double *)nullptr);
  } else {
    PHB_LAUNCH(c, k_lap<2>, grid, kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, gammaConst, gF, g0F, phi->cells0.p,
               m->nDev, theta, sign);
    if (m->nBCells)
      PHB_LAUNCH(c, k_lap_bnd<2>, gb, 256, 0, M, phi->dFaceType.p, e->vals.p, e->rhs.p, m->nLocal, gammaConst, gF,
                 g0F, phi->faces.p, phi->faces0.p, phi->cells0.p, m->nDev, theta, sign, tens);
  }
  return PHB_OK;
}

}  // extern "C"

namespace phb {
// uEqn_ of FractionalStep in one pass (k_momentum_fused) + the two boundary launches.  Returns 1 when the fused
// form does not apply (SYMMETRY patches carry a tensor block): the caller then assembles term by term.
int assemble_momentum_predictor(phb_eqn *e, phb_field *u, const phb_field *gradP, double gamma, double dt) {
  PHB_CHECK(check_pair(e, u, "assemble_momentum_predictor"));
  PHB_CHECK(check_pair(e, gradP, "assemble_momentum_predictor"));
  PHB_REQUIRE(e->nComp == 2 && u->nComp == 2 && gradP->nComp == 2, "assemble_momentum_predictor: vector equation expected");
  PHB_REQUIRE(u->hasOld, "assemble_momentum_predictor: u has no previous time step (savePreviousTimeStep)");
  PHB_REQUIRE(dt > 0., "assemble_momentum_predictor: dt must be positive");
  for (const BcEntry &b : u->bc)
    if (b.type == PHB_SYMMETRY) return 1;
  phb_mesh *m = e->m;
  phb_ctx *c = m->ctx;
  PHB_CHECK(phb::field_face_types(u));
  if (e->tens.p) PHB_CHECK(e->tens.zero(c->stream));
  e->hasTens = false;
  const MeshView M = view(m);
  const int grid = std::min(row_grid(c, m), c->numSMs * kFusedBlocksPerSM), gb = (m->nBCells + 255) / 256;
  PHB_LAUNCH(c, k_momentum_fused<2>, grid, kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, u->faces0.p, u->cells0.p,
             m->nDev, gradP->cells.p, gamma, dt);
  if (m->nBCells) {
    PHB_LAUNCH(c, (k_div_bnd<2, 0>), gb, 256, 0, M, u->dFaceType.p, e->vals.p, e->rhs.p, m->nLocal, u->faces.p,
               u->faces0.p, u->faces0.p, u->faces.p, u->faces0.p, u->faces0.p, u->cells0.p, m->nDev, 0., +1.);
    PHB_LAUNCH(c, k_lap_bnd<2>, gb, 256, 0, M, u->dFaceType.p, e->vals.p, e->rhs.p, m->nLocal, gamma,
               (const double *)nullptr, (const double *)nullptr, u->faces.p, u->faces0.p, u->cells0.p, m->nDev, 0.5,
               -1., (double *)nullptr);
  }
  return PHB_OK;
}
// pEqn_ of FractionalStep in one pass (k_pressure_fused) + the two boundary launches
int assemble_pressure_poisson(phb_eqn *e, phb_field *p, const phb_field *u, double gamma) {
  PHB_CHECK(check_pair(e, p, "assemble_pressure_poisson"));
  PHB_CHECK(check_pair(e, u, "assemble_pressure_poisson"));
  PHB_REQUIRE(e->nComp == 1 && p->nComp == 1 && u->nComp == 2, "assemble_pressure_poisson: scalar equation, vector flux field");
  phb_mesh *m = e->m;
  phb_ctx *c = m->ctx;
  PHB_CHECK(phb::field_face_types(p));
  if (e->tens.p) PHB_CHECK(e->tens.zero(c->stream));
  e->hasTens = false;
  const MeshView M = view(m);
  const int gb = (m->nBCells + 255) / 256;
  PHB_LAUNCH(c, k_pressure_fused, row_grid(c, m), kThreads, 0, M, e->vals.p, e->rhs.p, u->faces.p, gamma);
  if (m->nBCells) {
    PHB_LAUNCH(c, k_lap_bnd<1>, gb, 256, 0, M, p->dFaceType.p, e->vals.p, e->rhs.p, m->nLocal, gamma,
               (const double *)nullptr, (const double *)nullptr, p->faces.p, p->faces0.p, p->cells0.p, m->nDev, -1., +1.,
               (double *)nullptr);
    PHB_LAUNCH(c, k_flux_sum_bnd<0>, gb, 256, 0, M, u->faces.p, e->rhs.p, -1., 0., (const int *)nullptr);
  }
  return PHB_OK;
}
}  // namespace phb

extern "C" {
int phb_assemble_src(phb_eqn *e, const phb_field *f, double sign) {
  PHB_CHECK(check_pair(e, f, "phb_assemble_src"));
  PHB_REQUIRE(f->nComp == e->nComp, "phb_assemble_src: component mismatch");
  phb_mesh *m = e->m;
  const MeshView M = view(m);
  if (e->nComp == 1)
    PHB_LAUNCH(m->ctx, k_src<1>, row_grid(m->ctx, m), kThreads, 0, M, e->rhs.p, m->nLocal, f->cells.p, m->nDev, sign);
  else
    PHB_LAUNCH(m->ctx, k_src<2>, row_grid(m->ctx, m), kThreads, 0, M, e->rhs.p, m->nLocal, f->cells.p, m->nDev, sign);
  return PHB_OK;
}

static int assemble_src_div(phb_eqn *e, const phb_field *u, double sign, const int *mask) {
  PHB_CHECK(check_pair(e, u, "phb_assemble_src_div"));
  PHB_REQUIRE(e->nComp == 1 && u->nComp == 2, "phb_assemble_src_div: scalar equation, vector field");
  phb_mesh *m = e->m;
  const MeshView M = view(m);
  PHB_LAUNCH(m->ctx, k_flux_sum<0>, row_grid(m->ctx, m), kThreads, 0, M, u->faces.p, e->rhs.p, sign, 0., (const int *)mask);
  if (m->nBCells)
    PHB_LAUNCH(m->ctx, k_flux_sum_bnd<0>, (m->nBCells + 255) / 256, 256, 0, M, u->faces.p, e->rhs.p, sign, 0., (const int *)mask);
  return PHB_OK;
}
int phb_assemble_src_div(phb_eqn *e, const phb_field *u, double sign) { return assemble_src_div(e, u, sign, nullptr); }
// src::div(field, cells) (UD/Source.cpp:5-21): the listed cells only, the other rows get nothing
int phb_assemble_src_div_cells(phb_eqn *e, const phb_field *u, double sign, int nCells, const int *cells) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(e && cells && nCells >= 0, "phb_assemble_src_div_cells: bad argument");
  phb::DevBuf<int> mask;
  PHB_CHECK(cell_mask(e->m, nCells, cells, mask));
  PHB_CHECK(assemble_src_div(e, u, sign, mask.p));
  PHB_CUDA(cudaStreamSynchronize(e->m->ctx->stream));
  return PHB_OK;
  PHB_TRY_END
}

// src::laplacian (UD/Source.cpp:27-75): rhs += sign * sum_links c (phi_nb - phi_P), c = Gamma g_f; boundary links use
// the face value of phi whatever the patch type.  The field-Gamma overload of the reference (:50-75) sizes its result
// 2N and writes row indexMap->local(cell, 1) of a one-component index map -- an out-of-bounds read that crashes when
// the reference's own code is run (oracle/_ref) -- so only its evident intent exists here: the same scalar sum with
// Gamma_f taken from the field's face values.
int phb_assemble_src_laplacian(phb_eqn *e, double gammaConst, const phb_field *gam, const phb_field *phi, double sign) {
  PHB_CHECK(check_pair(e, phi, "phb_assemble_src_laplacian"));
  PHB_REQUIRE(phi->nComp == 1, "phb_assemble_src_laplacian: phi must be a scalar field");
  PHB_REQUIRE(!gam || (gam->nComp == 1 && gam->m == e->m), "phb_assemble_src_laplacian: gamma must be a scalar field");
  PHB_REQUIRE(e->nComp == 1, "phb_assemble_src_laplacian: scalar equation expected");
  phb_mesh *m = e->m;
  const MeshView M = view(m);
  const double *gF = gam ? gam->faces.p : nullptr;
  PHB_LAUNCH(m->ctx, k_src_lap, row_grid(m->ctx, m), kThreads, 0, M, e->rhs.p, gammaConst, gF, phi->cells.p, sign);
  if (m->nBCells)
    PHB_LAUNCH(m->ctx, k_src_lap_bnd, (m->nBCells + 255) / 256, 256, 0, M, e->rhs.p, gammaConst, gF, phi->cells.p,
               phi->faces.p, sign);
  return PHB_OK;
}

int phb_eqn_scale_rows(phb_eqn *e, const phb_field *rho) {
  PHB_CHECK(check_pair(e, rho, "phb_eqn_scale_rows"));
  PHB_REQUIRE(rho->nComp == 1, "phb_eqn_scale_rows: rho must be a scalar field");
  phb_mesh *m = e->m;
  const MeshView M = view(m);
  if (e->nComp == 1)
    PHB_LAUNCH(m->ctx, k_scale_rows<1>, row_grid(m->ctx, m), kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, rho->cells.p);
  else
    PHB_LAUNCH(m->ctx, k_scale_rows<2>, row_grid(m->ctx, m), kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, rho->cells.p);
  if (e->hasTens)
    PHB_LAUNCH(m->ctx, k_scale_tens, flat_grid(m->ctx, 4ll * m->nLocal), kThreads, 0, m->nLocal, rho->cells.p, e->tens.p);
  return PHB_OK;
}

int phb_eqn_relax(phb_eqn *e, const phb_field *phi, double omega) {
  PHB_CHECK(check_pair(e, phi, "phb_eqn_relax"));
  PHB_REQUIRE(phi->nComp == e->nComp && omega > 0., "phb_eqn_relax: bad argument");
  phb_mesh *m = e->m;
  const MeshView M = view(m);
  if (e->nComp == 1)
    PHB_LAUNCH(m->ctx, k_relax<1>, row_grid(m->ctx, m), kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, phi->cells.p, m->nDev, omega);
  else
    PHB_LAUNCH(m->ctx, k_relax<2>, row_grid(m->ctx, m), kThreads, 0, M, e->vals.p, e->rhs.p, m->nLocal, phi->cells.p, m->nDev, omega);
  return PHB_OK;
}

long long phb_eqn_export_csr(const phb_eqn *e, int layout, int *rowPtr, int *colInd, double *vals, double *rhs) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(e && layout >= 0 && layout <= 2, "phb_eqn_export_csr: bad argument");
  phb_mesh *m = e->m;
  const SellPattern &S = m->sell;
  const int nL = m->nLocal, nc = e->nComp;
  PHB_REQUIRE(layout == 0 || nc == 1, "phb_eqn_export_csr: padded layouts are scalar-only");
  std::vector<double> hv((size_t)S.nSlots), hr((size_t)nc * nL), ht;
  if (e->hasTens) {
    ht.resize(4 * (size_t)nL);
    PHB_CUDA(cudaMemcpyAsync(ht.data(), e->tens.p, ht.size() * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  }
  PHB_CUDA(cudaMemcpyAsync(hv.data(), e->vals.p, hv.size() * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  PHB_CUDA(cudaMemcpyAsync(hr.data(), e->rhs.p, hr.size() * sizeof(double), cudaMemcpyDeviceToHost, m->ctx->stream));
  PHB_CUDA(cudaStreamSynchronize(m->ctx->stream));
  auto slot = [&](int d, int k) { return (size_t)S.hSliceOff[d >> 5] + (size_t)k * 32 + (d & 31); };
  // global column of entry k of device row d: canonical CSR holds globalRow ids in the same order
  long long nnz = 0;
  std::vector<int> rp, ci;
  std::vector<double> va;
  rp.push_back(0);
  // vector IndexMap (UE/IndexMap.cpp:29-36): rows k*nLocal + local, cols 2*offset + k*nLocal + local
  PHB_REQUIRE(nc == 1 || m->nProcs == 1, "phb_eqn_export_csr: vector export is single-process only");
  for (int comp = 0; comp < nc; ++comp)
    for (int d = 0; d < nL; ++d) {
      const int len = S.hRowLen[d];
      const int base = m->rowPtr[d];
      auto col = [&](int k) { return m->colInd[base + k] + comp * nL; };
      if (layout == 0) {
        for (int k = 0; k < len; ++k) {
          double v = hv[slot(d, k)];
          if (k == 0 && !ht.empty()) v += ht[(comp == 0 ? 0 : 3) * (size_t)nL + d];   // xx / yy join the diagonal
          if (k > 0 && v == 0.) continue;  // exact zeros are dropped by operator+= (M/CrsEquation.cpp:196-206)
          ci.push_back(col(k)); va.push_back(v);
        }
        if (!ht.empty()) {   // xy / yx: a new column (the cell's other component), appended by the merge, only when non-zero
          const double x = ht[(comp == 0 ? 1 : 2) * (size_t)nL + d];
          if (x != 0.) { ci.push_back(m->colInd[base] + (1 - comp) * nL); va.push_back(x); }
        }
      } else {
        // ELL-5: fv::laplacian(Scalar,phi) inserts nb0 first, then P (layout 1); the Field overload P first (layout 2)
        std::vector<int> order;
        if (layout == 1 && len > 1) { order.push_back(1); order.push_back(0); for (int k = 2; k < len; ++k) order.push_back(k); }
        else for (int k = 0; k < len; ++k) order.push_back(k);
        for (int k : order) { ci.push_back(col(k)); va.push_back(hv[slot(d, k)]); }
        for (int k = len; k < 5; ++k) { ci.push_back(-1); va.push_back(0.); }
      }
      rp.push_back((int)ci.size());
    }
  nnz = (long long)ci.size();
  if (rowPtr) std::copy(rp.begin(), rp.end(), rowPtr);
  if (colInd) std::copy(ci.begin(), ci.end(), colInd);
  if (vals) std::copy(va.begin(), va.end(), vals);
  if (rhs) std::copy(hr.begin(), hr.end(), rhs);
  return nnz;
  PHB_TRY_END
}

int phb_eqn_solve(phb_eqn *e, phb_solver *s, phb_field *phi, int warmStart, int *iters, double *relres) {
  return phb::eqn_solve_tagged(e, s, phi, warmStart, 0ull, iters, relres);
}
}  // extern "C"

namespace phb {
// `tag` != 0: the caller vouches that equal tags mean equal coefficients (the one-pass assembly of FractionalStep: the
// matrix is a function of dt, the diffusivity and the boundary types only) -- the multigrid preconditioner then skips
// its per-solve comparison of the coefficients with the ones it was built from (one sweep over both + a host sync)
int eqn_solve_tagged(phb_eqn *e, phb_solver *s, phb_field *phi, int warmStart, unsigned long long tag, int *iters,
                     double *relres) {
  PHB_TRY_BEGIN
  PHB_CHECK(check_pair(e, phi, "phb_eqn_solve"));
  PHB_REQUIRE(s && phi->nComp == e->nComp, "phb_eqn_solve: bad argument");
  phb_mesh *m = e->m;
  phb_ctx *c = m->ctx;
  PHB_REQUIRE(s->ctx == c, "phb_eqn_solve: solver belongs to another context");
  PHB_CHECK(phb::solver_bind(s, &m->sell, e->vals.p, e->nComp, c->nProcs > 1 ? m : nullptr));
  s->tens = e->hasTens ? e->tens.p : nullptr;
  const int n = m->nLocal, nc = e->nComp, ld = s->ld;
  PHB_LAUNCH(c, k_neg_copy, flat_grid(c, (long long)n * nc), kThreads, 0, n, nc, n, ld, e->rhs.p, s->b.p);
  if (warmStart)
    PHB_LAUNCH(c, k_copy2d, flat_grid(c, (long long)n * nc), kThreads, 0, n, nc, (long long)m->nDev, (long long)ld,
               phi->cells.p, s->x.p);
  else
    PHB_CUDA(cudaMemsetAsync(s->x.p, 0, (size_t)ld * nc * sizeof(double), c->stream));
  s->valsTag = tag;
  int rc = phb::solver_run(s, iters, relres);
  s->valsTag = 0ull;
  if (rc != PHB_OK) return rc;
  // mapFromSparseSolver: x -> field cells (UE/ScalarFiniteVolumeEquation.cpp:59-64)
  PHB_LAUNCH(c, k_copy2d, flat_grid(c, (long long)n * nc), kThreads, 0, n, nc, (long long)ld, (long long)m->nDev,
             s->x.p, phi->cells.p);
  return PHB_OK;
  PHB_TRY_END
}
}  // namespace phb
