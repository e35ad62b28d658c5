// multiphase.cu -- device-resident FractionalStepMultiphase::solve (config 4: rising bubble, VOF + surface tension).
//
// Reference (under /root/reference/src/2D/Unstructured):
//   Solvers/FractionalStepMultiphase.cpp:52-218        initialize, solve, solveGammaEqn / UEqn / PEqn,
//                                                      correctVelocity, updateProperties
//   FiniteVolume/Field/VectorFiniteVolumeField.cpp:28-50   faceToCell(cellWeight, faceWeight, cells)
//   FiniteVolume/Multiphase/SurfaceTensionForce.cpp:50-118 interface normals with contact-angle boundary faces,
//                                                      smoothed gamma
//   FiniteVolume/Multiphase/SurfaceTensionForceSmoothingKernel.cpp:3-59  kernel support + weights (pow8 / pow6 / peskin)
//   FiniteVolume/Multiphase/Celeste.cpp:10-108, CelesteStencil.cpp:9-137 least-squares gradient / curvature stencils
// The CICSAM advection, the operators and the three solves are the existing kernels (cicsam.cu, fv.cu, solver.cu).
//
// CELESTE on the device: the reference keeps, per cell, a 5 x m pseudo-inverse of the Taylor matrix of its stencil
// (face + node neighbours and nearby boundary faces) and multiplies it with the stencil differences every step.
// Only the two gradient rows of that pseudo-inverse are ever used, so the host setup stores, per stencil entry,
// the two coefficients (already divided by the weight s_k): grad phi = sum_k c_k (phi_k - phi_P), div n = sum_k
// (cx_k (nx_k - nx_P) + cy_k (ny_k - ny_P)) -- a gather over a CSR stencil list, one lane per cell.
// Single rank only (the reference exchanges n, kappa, sg, fst between ranks; not built here).
#include <algorithm>
#include <cmath>
#include <thread>

#include "fv.cuh"
#include "kernels.cuh"
#include "solver.cuh"

using namespace phb;

struct phb_multiphase {
  phb_mesh *m = nullptr;
  double rho1 = 1., rho2 = 1., mu1 = 1., mu2 = 1., sigma = 0., gx = 0., gy = 0., eps = 1e-8, radius = 0.;
  int kernelType = 2;   // 0 peskin, 1 pow6, 2 pow8
  std::vector<double> patchTheta;
  phb_field *u = nullptr, *p = nullptr, *gradP = nullptr, *gamma = nullptr, *gradGamma = nullptr, *rho = nullptr,
            *mu = nullptr, *beta = nullptr, *sg = nullptr, *fst = nullptr, *kappa = nullptr, *gammaTilde = nullptr,
            *gradGammaTilde = nullptr, *n = nullptr, *gradRho = nullptr, *force = nullptr, *dtRho = nullptr;
  phb_eqn *gammaEqn = nullptr, *uEqn = nullptr, *pEqn = nullptr;
  phb_solver *gammaSolver = nullptr, *uSolver = nullptr, *pSolver = nullptr;
  // CELESTE
  bool built = false;
  DevBuf<int> kPtr, kCell, stPtr, stIdx;
  DevBuf<double> kW, stKappa, stGrad, faceVW, faceC, bfTheta;
  DevBuf<double> scratch, partials, out;
  DevBuf<unsigned> ticket;
  bool warmStart = true;
};

namespace {
constexpr int kThreads = 256;

struct View {
  const int *sliceOff, *col, *linkFace;
  int nRows, nSlices, nDev, nFaces, nBCells, nBFaces;
  const double *vol, *fSx, *fSy, *fQx, *fQy;
  const int *fL, *fR, *bcCell, *bcPtr, *bcFace, *bfFace, *bfCell;
};
View view(const phb_mesh *m) {
  View v;
  v.sliceOff = m->sell.sliceOff.p; v.col = m->sell.col.p; v.linkFace = m->dLinkFace.p;
  v.nRows = m->sell.nRows; v.nSlices = m->sell.nSlices; v.nDev = m->nDev; v.nFaces = m->nFaces;
  v.nBCells = m->nBCells; v.nBFaces = m->nBFaces;
  v.vol = m->dVol.p; v.fSx = m->dFSx.p; v.fSy = m->dFSy.p; v.fQx = m->dFQx.p; v.fQy = m->dFQy.p;
  v.fL = m->dFL.p; v.fR = m->dFR.p;
  v.bcCell = m->dBcCell.p; v.bcPtr = m->dBcPtr.p; v.bcFace = m->dBcFace.p; v.bfFace = m->dBfFace.p; v.bfCell = m->dBfCell.p;
  return v;
}
int grid_flat(const phb_ctx *c, long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + kThreads - 1) / kThreads, (long long)c->numSMs * 8));
}
__device__ __forceinline__ double clamp01(double v) { return fmax(fmin(v, 1.), 0.); }

// rho = rho1 + clamp(gamma)(rho2 - rho1); mu = rho / (rho1/mu1 + clamp(gamma)(rho2/mu2 - rho1/mu1))   (:175-205)
__global__ void k_properties(long long n, const double *__restrict__ gam, double rho1, double rho2, double mu1, double mu2,
                             double *__restrict__ rho, double *__restrict__ mu) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const double g = clamp01(gam[i]);
    const double r = rho1 + g * (rho2 - rho1);
    rho[i] = r;
    mu[i] = r / (rho1 / mu1 + g * (rho2 / mu2 - rho1 / mu1));
  }
}
// sg_f = dot(g, -c_f) gradRho_f   (:186-188)
__global__ void k_sg_faces(int nF, const double *__restrict__ fc, double gx, double gy, const double *__restrict__ gradRhoF,
                           double *__restrict__ sgF) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
    const double s = -(gx * fc[f] + gy * fc[(size_t)nF + f]);
    sgF[f] = s * gradRhoF[f];
    sgF[(size_t)nF + f] = s * gradRhoF[(size_t)nF + f];
  }
}
// fst_f = sigma kappa_f gradGamma_f   (Celeste.cpp:18-19)
__global__ void k_fst_faces(int nF, double sigma, const double *__restrict__ kappaF, const double *__restrict__ gradGF,
                            double *__restrict__ fstF) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nF; f += gridDim.x * blockDim.x) {
    fstF[f] = sigma * kappaF[f] * gradGF[f];
    fstF[(size_t)nF + f] = sigma * kappaF[f] * gradGF[(size_t)nF + f];
  }
}
// VectorFiniteVolumeField::faceToCell(cellWeight, faceWeight, cells): c_P = w_P sum_f (v_f (.) |S_f| / w_f) / sum_f |S_f|
__global__ void k_face_to_cell_weighted(View M, const double *__restrict__ vF, const double *__restrict__ cellW,
                                        const double *__restrict__ faceW, double *__restrict__ vC) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int slice = blockIdx.x * wpb + (threadIdx.x >> 5); slice < M.nSlices; slice += gridDim.x * wpb) {
    const int off = M.sliceOff[slice], wdt = (M.sliceOff[slice + 1] - off) >> 5, row = slice * 32 + lane;
    if (row >= M.nRows) continue;
    double tx = 0., ty = 0., sx = 0., sy = 0.;
    for (int k = 1; k < wdt; ++k) {
      const int lf = M.linkFace[(size_t)off + (size_t)k * 32 + lane];
      if (lf < 0) continue;
      const int f = lf >> 1;
      const double ax = fabs(M.fSx[f]), ay = fabs(M.fSy[f]), w = faceW[f];
      tx += vF[f] * ax / w; ty += vF[(size_t)M.nFaces + f] * ay / w;
      sx += ax; sy += ay;
    }
    const double w = cellW[row];    // cells with boundary links are redone, boundary faces included, by the second launch
    vC[row] = w * tx / sx; vC[(size_t)M.nDev + row] = w * ty / sy;
  }
}
// boundary cells: the full sum again with the boundary links included (tiny launch)
__global__ void k_face_to_cell_weighted_bnd(View M, const double *__restrict__ vF, const double *__restrict__ cellW,
                                            const double *__restrict__ faceW, double *__restrict__ vC) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  const int off = M.sliceOff[row >> 5], wdt = (M.sliceOff[(row >> 5) + 1] - off) >> 5;
  double tx = 0., ty = 0., sx = 0., sy = 0.;
  for (int k = 1; k < wdt; ++k) {
    const int lf = M.linkFace[(size_t)off + (size_t)k * 32 + (row & 31)];
    if (lf < 0) continue;
    const int f = lf >> 1;
    const double ax = fabs(M.fSx[f]), ay = fabs(M.fSy[f]), w = faceW[f];
    tx += vF[f] * ax / w; ty += vF[(size_t)M.nFaces + f] * ay / w;
    sx += ax; sy += ay;
  }
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    const double ax = fabs(M.fSx[f]), ay = fabs(M.fSy[f]), w = faceW[f];
    tx += vF[f] * ax / w; ty += vF[(size_t)M.nFaces + f] * ay / w;
    sx += ax; sy += ay;
  }
  const double w = cellW[row];
  vC[row] = w * tx / sx; vC[(size_t)M.nDev + row] = w * ty / sy;
}

// gammaTilde_P = sum_k w_k gamma_k   (SmoothingKernel::eval; the weights hold kernel * volume * A_)
__global__ void k_smooth(int n, const int *__restrict__ kPtr, const int *__restrict__ kCell, const double *__restrict__ kW,
                         const double *__restrict__ gam, double *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.;
    for (int k = kPtr[i]; k < kPtr[i + 1]; ++k) s += kW[k] * gam[kCell[k]];
    out[i] = s;
  }
}
// Celeste::Stencil::grad: sum_k c_k (phi_k - phi_P), entries = cells (idx >= 0) or boundary faces (~idx)
__global__ void k_stencil_grad(int n, int nDev, const int *__restrict__ stPtr, const int *__restrict__ stIdx,
                               const double *__restrict__ coef, const double *__restrict__ phiC,
                               const double *__restrict__ phiF, double *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double p0 = phiC[i];
    double gx = 0., gy = 0.;
    for (int k = stPtr[i]; k < stPtr[i + 1]; ++k) {
      const int id = stIdx[k];
      const double d = (id >= 0 ? phiC[id] : phiF[~id]) - p0;
      gx += coef[2 * (size_t)k] * d;
      gy += coef[2 * (size_t)k + 1] * d;
    }
    out[i] = gx;
    out[(size_t)nDev + i] = gy;
  }
}
// n_P = -gradGammaTilde / |.| where |.|^2 >= eps^2, else 0   (SurfaceTensionForce.cpp:55-58)
__global__ void k_normals(int n, int nDev, double eps2, const double *__restrict__ g, double *__restrict__ nC) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = g[i], y = g[(size_t)nDev + i], m2 = x * x + y * y;
    if (m2 >= eps2) {
      const double inv = 1. / sqrt(m2);
      nC[i] = -x * inv; nC[(size_t)nDev + i] = -y * inv;
    } else {
      nC[i] = 0.; nC[(size_t)nDev + i] = 0.;
    }
  }
}
// boundary faces from the contact-line orientation   (SurfaceTensionForce.cpp:62-81)
__global__ void k_normals_bnd(View M, const double *__restrict__ theta, const double *__restrict__ nC, double *__restrict__ nF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBFaces) return;
  const int f = M.bfFace[i], l = M.bfCell[i];
  const double nx = nC[l], ny = nC[(size_t)M.nDev + l];
  if (nx * nx + ny * ny == 0.) { nF[f] = 0.; nF[(size_t)M.nFaces + f] = 0.; return; }
  const double sm = sqrt(M.fSx[f] * M.fSx[f] + M.fSy[f] * M.fSy[f]);
  const double nsx = -M.fSx[f] / sm, nsy = -M.fSy[f] / sm;     // S_f of a boundary face points out of its cell
  const double d = nx * nsx + ny * nsy, nn = nsx * nsx + nsy * nsy;
  double tx = nx - d * nsx / nn, ty = ny - d * nsy / nn;        // tangentialComponent(ns)
  const double tm = sqrt(tx * tx + ty * ty);
  tx /= tm; ty /= tm;
  if (isnan(tx) || isnan(ty)) { nF[f] = nx; nF[(size_t)M.nFaces + f] = ny; return; }
  const double th = theta[i];
  nF[f] = nsx * cos(th) + tx * sin(th);
  nF[(size_t)M.nFaces + f] = nsy * cos(th) + ty * sin(th);
}
// kappa_P = div n over the stencil where n is non-zero on the cell and all its face / node neighbours, else 0
// (Celeste.cpp:63-84); `nbOnly` entries = the first nbCells[i] stencil entries are exactly those neighbours
__global__ void k_curvature(int n, int nDev, int nFaces, const int *__restrict__ stPtr, const int *__restrict__ stIdx,
                            const double *__restrict__ coef, const double *__restrict__ nC, const double *__restrict__ nF,
                            double *__restrict__ kappa) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x0 = nC[i], y0 = nC[(size_t)nDev + i];
    bool ok = x0 * x0 + y0 * y0 != 0.;
    double s = 0.;
    for (int k = stPtr[i]; k < stPtr[i + 1]; ++k) {
      const int id = stIdx[k];
      double x, y;
      if (id >= 0) {
        x = nC[id]; y = nC[(size_t)nDev + id];
        if (x * x + y * y == 0.) ok = false;
      } else {
        x = nF[~id]; y = nF[(size_t)nFaces + ~id];
      }
      s += coef[2 * (size_t)k] * (x - x0) + coef[2 * (size_t)k + 1] * (y - y0);
    }
    kappa[i] = ok ? s : 0.;
  }
}
// face curvature "according to Afkhami 2007"   (Celeste.cpp:88-107); boundary faces keep their value when n_l = 0
__global__ void k_curvature_faces(View M, const double *__restrict__ vw, const double *__restrict__ nC,
                                  const double *__restrict__ kC, double *__restrict__ kF) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < M.nFaces; f += gridDim.x * blockDim.x) {
    const int l = M.fL[f], r = M.fR[f];
    const bool hl = nC[l] * nC[l] + nC[(size_t)M.nDev + l] * nC[(size_t)M.nDev + l] != 0.;
    if (r < 0) {
      if (hl) kF[f] = kC[l];
      continue;
    }
    const bool hr = nC[r] * nC[r] + nC[(size_t)M.nDev + r] * nC[(size_t)M.nDev + r] != 0.;
    const double g = vw[f];
    kF[f] = hl && hr ? g * kC[l] + (1. - g) * kC[r] : hl ? kC[l] : hr ? kC[r] : 0.;
  }
}
// force = fst + sg - gradP on the owned cells   (:112)
__global__ void k_force(int n, int nDev, const double *__restrict__ a, const double *__restrict__ b, const double *__restrict__ c,
                        double *__restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += gridDim.x * blockDim.x) {
    const size_t k = (size_t)(i / n) * nDev + (i % n);
    out[k] = a[k] + b[k] - c[k];
  }
}
// u_P += s dt / rho_P gradP_P   (:116-117, :166-167)
__global__ void k_axpy_over_rho(int n, int nDev, double a, const double *__restrict__ rho, const double *__restrict__ g,
                                double *__restrict__ u) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += gridDim.x * blockDim.x) {
    const int c = i / n, r = i % n;
    u[(size_t)c * nDev + r] += a / rho[r] * g[(size_t)c * nDev + r];
  }
}
__global__ void k_axpy_over_rho_faces(int nF, double a, const double *__restrict__ rhoF, const double *__restrict__ g,
                                      double *__restrict__ u) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * nF; i += gridDim.x * blockDim.x)
    u[i] += a / rhoF[i % nF] * g[i];
}
// momentum-weighted face velocity   (:121-130)
__global__ void k_momentum_faces(View M, const double *__restrict__ vw, double dt, const double *__restrict__ uC,
                                 const double *__restrict__ rhoC, const double *__restrict__ rhoF,
                                 const double *__restrict__ fC, const double *__restrict__ fF, double *__restrict__ uF) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < M.nFaces; f += gridDim.x * blockDim.x) {
    const int l = M.fL[f], r = M.fR[f];
    if (r < 0) continue;
    const double g = vw[f];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const size_t oc = (size_t)c * M.nDev, of = (size_t)c * M.nFaces;
      uF[of + f] = g * (uC[oc + l] - dt / rhoC[l] * fC[oc + l]) + (1. - g) * (uC[oc + r] - dt / rhoC[r] * fC[oc + r]) +
                   dt / rhoF[f] * fF[of + f];
    }
  }
}
// boundary patches   (:132-147)
__global__ void k_momentum_faces_bnd(View M, const int *__restrict__ bfType, double dt, const double *__restrict__ uC,
                                     const double *__restrict__ rhoC, const double *__restrict__ rhoF,
                                     const double *__restrict__ fC, const double *__restrict__ fF, double *__restrict__ uF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBFaces) return;
  const int f = M.bfFace[i], l = M.bfCell[i], t = bfType[i];
  if (t == PHB_NORMAL_GRADIENT) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const size_t oc = (size_t)c * M.nDev, of = (size_t)c * M.nFaces;
      uF[of + f] = uC[oc + l] - dt / rhoC[l] * fC[oc + l] + dt / rhoF[f] * fF[of + f];
    }
  } else if (t == PHB_SYMMETRY) {
    const double nx = M.fSx[f], ny = M.fSy[f], ux = uC[l], uy = uC[(size_t)M.nDev + l];
    const double d = ux * nx + uy * ny, mm = nx * nx + ny * ny;
    uF[f] = ux - d * nx / mm;
    uF[(size_t)M.nFaces + f] = uy - d * ny / mm;
  }
}
// out = a / x over cells and faces   (timeStep / rho_, :153)
__global__ void k_recip(long long n, double a, const double *__restrict__ x, double *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a / x[i];
}
__global__ void k_sum_fields(long long n, const double *__restrict__ a, const double *__restrict__ b, double *__restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = a[i] + b[i];
}

// ---- host setup -----------------------------------------------------------------------------------------------
double kernel_value(int type, double eps, double dx, double dy) {
  auto kcos = [eps](double x) { return x < eps ? eps * (1. + std::cos(M_PI * x / eps)) : 0.; };
  if (type == 0) return kcos(dx) * kcos(dy);
  const double r = std::sqrt(dx * dx + dy * dy);
  if (!(r < eps)) return 0.;
  const double b = eps * eps - r * r;
  return type == 1 ? b * b * b : b * b * b * b;
}

// rows 3 and 4 of the pseudo-inverse of the m x 5 Taylor matrix (LAPACKE_dgels on the identity in the reference,
// Math/Matrix.cpp:229-235).  m >= 5: QR of the column-scaled matrix (the columns differ by a factor h), twice
// re-orthogonalised Gram-Schmidt; m < 5: minimum-norm solution A^T (A A^T)^-1.
void pinv_rows(const std::vector<double> &A, int m, std::vector<double> &row3, std::vector<double> &row4) {
  row3.assign(m, 0.); row4.assign(m, 0.);
  if (m >= 5) {
    double d[5], R[5][5] = {{0.}};
    std::vector<double> q(5 * (size_t)m);
    for (int j = 0; j < 5; ++j) {
      double s = 0.;
      for (int i = 0; i < m; ++i) s += A[i * 5 + j] * A[i * 5 + j];
      d[j] = s > 0. ? std::sqrt(s) : 1.;
      for (int i = 0; i < m; ++i) q[(size_t)j * m + i] = A[i * 5 + j] / d[j];
    }
    for (int j = 0; j < 5; ++j) {
      double *qj = &q[(size_t)j * m];
      for (int pass = 0; pass < 2; ++pass)
        for (int k = 0; k < j; ++k) {
          const double *qk = &q[(size_t)k * m];
          double r = 0.;
          for (int i = 0; i < m; ++i) r += qk[i] * qj[i];
          R[k][j] += r;
          for (int i = 0; i < m; ++i) qj[i] -= r * qk[i];
        }
      double nn = 0.;
      for (int i = 0; i < m; ++i) nn += qj[i] * qj[i];
      nn = std::sqrt(nn);
      R[j][j] = nn;
      if (nn > 0.) for (int i = 0; i < m; ++i) qj[i] /= nn;
    }
    // row t of pinv(As) = (Q R^-T e_t)^T : solve R^T y = e_t, row = Q y; then unscale by 1 / d_t
    for (int t = 3; t <= 4; ++t) {
      double y[5] = {0., 0., 0., 0., 0.};
      for (int i = 0; i < 5; ++i) {
        double s = i == t ? 1. : 0.;
        for (int k = 0; k < i; ++k) s -= R[k][i] * y[k];
        y[i] = R[i][i] != 0. ? s / R[i][i] : 0.;
      }
      std::vector<double> &row = t == 3 ? row3 : row4;
      for (int i = 0; i < m; ++i) {
        double s = 0.;
        for (int k = 0; k < 5; ++k) s += q[(size_t)k * m + i] * y[k];
        row[i] = s / d[t];
      }
    }
    return;
  }
  // m < 5: G = A A^T (m x m), pinv = A^T G^-1
  long double G[4][8];
  for (int i = 0; i < m; ++i) {
    for (int j = 0; j < m; ++j) {
      long double s = 0.;
      for (int k = 0; k < 5; ++k) s += (long double)A[i * 5 + k] * A[j * 5 + k];
      G[i][j] = s;
    }
    for (int j = 0; j < m; ++j) G[i][m + j] = i == j ? 1. : 0.;
  }
  for (int c = 0; c < m; ++c) {
    int piv = c;
    for (int r = c + 1; r < m; ++r) if (fabsl(G[r][c]) > fabsl(G[piv][c])) piv = r;
    if (G[piv][c] == 0.) return;
    if (piv != c) for (int j = 0; j < 2 * m; ++j) std::swap(G[c][j], G[piv][j]);
    const long double inv = 1. / G[c][c];
    for (int j = 0; j < 2 * m; ++j) G[c][j] *= inv;
    for (int r = 0; r < m; ++r) {
      if (r == c) continue;
      const long double l = G[r][c];
      for (int j = 0; j < 2 * m; ++j) G[r][j] -= l * G[c][j];
    }
  }
  for (int i = 0; i < m; ++i) {
    long double s3 = 0., s4 = 0.;
    for (int k = 0; k < m; ++k) { s3 += (long double)A[k * 5 + 3] * G[k][m + i]; s4 += (long double)A[k * 5 + 4] * G[k][m + i]; }
    row3[i] = (double)s3; row4[i] = (double)s4;
  }
}

int build_celeste(phb_multiphase *mp) {
  phb_mesh *m = mp->m;
  phb_ctx *c = m->ctx;
  const int N = m->nCells, F = m->nFaces;
  // ---- stencils in the reference's order (CelesteStencil.cpp:9-33)
  std::vector<int> stPtr(N + 1, 0);
  for (int i = 0; i < N; ++i) {
    int cnt = (m->ilPtr[i + 1] - m->ilPtr[i]) + (m->dlPtr[i + 1] - m->dlPtr[i]) + (m->blPtr[i + 1] - m->blPtr[i]);
    if (m->blPtr[i + 1] > m->blPtr[i])
      for (int j = m->ilPtr[i]; j < m->ilPtr[i + 1]; ++j) cnt += m->blPtr[m->ilCell[j] + 1] - m->blPtr[m->ilCell[j]];
    stPtr[i + 1] = stPtr[i] + cnt;
  }
  std::vector<int> stIdx(stPtr[N]);
  std::vector<double> cK(2 * (size_t)stPtr[N]), cG(2 * (size_t)stPtr[N]);
  const int T = std::max(1, std::min(16, (int)std::thread::hardware_concurrency()));
  std::vector<std::thread> th;
  std::vector<int> bad(T, 0);
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      std::vector<int> cells, faces;
      std::vector<double> A, r3, r4;
      for (int i = (int)((long long)N * t / T); i < (int)((long long)N * (t + 1) / T); ++i) {
        cells.clear(); faces.clear();
        const bool onBoundary = m->blPtr[i + 1] > m->blPtr[i];
        for (int j = m->ilPtr[i]; j < m->ilPtr[i + 1]; ++j) {
          const int nb = m->ilCell[j];
          cells.push_back(nb);
          if (onBoundary)
            for (int q = m->blPtr[nb]; q < m->blPtr[nb + 1]; ++q) faces.push_back(m->blFace[q]);
        }
        for (int j = m->dlPtr[i]; j < m->dlPtr[i + 1]; ++j) cells.push_back(m->dlCell[j]);
        for (int q = m->blPtr[i]; q < m->blPtr[i + 1]; ++q) faces.push_back(m->blFace[q]);
        const int mm = (int)(cells.size() + faces.size());
        if (mm > 32 || mm != stPtr[i + 1] - stPtr[i]) { bad[t] = 1; continue; }
        for (int pass = 0; pass < 2; ++pass) {   // 0: curvature stencil (unweighted), 1: gradient stencil (weighted)
          A.assign((size_t)mm * 5, 0.);
          std::vector<double> s(mm, 1.);
          for (int k = 0; k < mm; ++k) {
            const bool isCell = k < (int)cells.size();
            const int id = isCell ? cells[k] : faces[k - cells.size()];
            const double rx = (isCell ? m->cCx[id] : m->fCx[id]) - m->cCx[i], ry = (isCell ? m->cCy[id] : m->fCy[id]) - m->cCy[i];
            if (pass == 1) s[k] = rx * rx + ry * ry;
            A[k * 5 + 0] = rx * rx / 2. / s[k]; A[k * 5 + 1] = ry * ry / 2. / s[k]; A[k * 5 + 2] = rx * ry / s[k];
            A[k * 5 + 3] = rx / s[k]; A[k * 5 + 4] = ry / s[k];
          }
          pinv_rows(A, mm, r3, r4);
          std::vector<double> &dst = pass == 0 ? cK : cG;
          for (int k = 0; k < mm; ++k) {
            dst[2 * ((size_t)stPtr[i] + k)] = r3[k] / s[k];
            dst[2 * ((size_t)stPtr[i] + k) + 1] = r4[k] / s[k];
          }
        }
        for (int k = 0; k < mm; ++k)
          stIdx[stPtr[i] + k] = k < (int)cells.size() ? m->cell2dev[cells[k]] : ~faces[k - cells.size()];
      }
    });
  for (auto &x : th) x.join();
  for (int b : bad) PHB_REQUIRE(!b, "multiphase: a CELESTE stencil holds more than 32 entries");
  // ---- smoothing kernels: cells with |c_k - c_P| < radius (Circle::isInside), uniform-grid search
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  for (int i = 0; i < N; ++i) { x0 = std::min(x0, m->cCx[i]); x1 = std::max(x1, m->cCx[i]); y0 = std::min(y0, m->cCy[i]); y1 = std::max(y1, m->cCy[i]); }
  const double R = mp->radius, cell = std::max(R, 1e-300);
  const int gx = (int)std::min<double>(4096., std::floor((x1 - x0) / cell) + 1.), gy = (int)std::min<double>(4096., std::floor((y1 - y0) / cell) + 1.);
  const double hx = (x1 - x0) / gx + 1e-300, hy = (y1 - y0) / gy + 1e-300;
  auto bx = [&](double x) { return std::min(gx - 1, std::max(0, (int)((x - x0) / hx))); };
  auto by = [&](double y) { return std::min(gy - 1, std::max(0, (int)((y - y0) / hy))); };
  std::vector<int> bPtr((size_t)gx * gy + 1, 0), bCell(N);
  for (int i = 0; i < N; ++i) bPtr[(size_t)by(m->cCy[i]) * gx + bx(m->cCx[i]) + 1]++;
  for (size_t b = 0; b < (size_t)gx * gy; ++b) bPtr[b + 1] += bPtr[b];
  {
    std::vector<int> fill(bPtr.begin(), bPtr.end() - 1);
    for (int i = 0; i < N; ++i) bCell[fill[(size_t)by(m->cCy[i]) * gx + bx(m->cCx[i])]++] = i;
  }
  const int rx = (int)std::ceil(R / hx), ry = (int)std::ceil(R / hy);
  std::vector<int> kCnt(N + 1, 0);
  std::vector<std::vector<int>> kc(T);
  std::vector<std::vector<double>> kw(T);
  th.clear();
  for (int t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      for (int i = (int)((long long)N * t / T); i < (int)((long long)N * (t + 1) / T); ++i) {
        const size_t start = kc[t].size();
        double A = 0.;
        const int ci = bx(m->cCx[i]), cj = by(m->cCy[i]);
        for (int j = std::max(0, cj - ry); j <= std::min(gy - 1, cj + ry); ++j)
          for (int a = std::max(0, ci - rx); a <= std::min(gx - 1, ci + rx); ++a)
            for (int q = bPtr[(size_t)j * gx + a]; q < bPtr[(size_t)j * gx + a + 1]; ++q) {
              const int k = bCell[q];
              const double dx = m->cCx[k] - m->cCx[i], dy = m->cCy[k] - m->cCy[i];
              if (!(dx * dx + dy * dy < R * R)) continue;
              const double w = kernel_value(mp->kernelType, R, dx, dy) * m->vol[k];
              kc[t].push_back(m->cell2dev[k]);
              kw[t].push_back(w);
              A += w;
            }
        for (size_t q = start; q < kw[t].size(); ++q) kw[t][q] /= A;
        kCnt[i + 1] = (int)(kc[t].size() - start);
      }
    });
  for (auto &x : th) x.join();
  for (int i = 0; i < N; ++i) kCnt[i + 1] += kCnt[i];
  std::vector<int> kCell(kCnt[N]);
  std::vector<double> kW(kCnt[N]);
  {
    size_t o = 0;
    for (int t = 0; t < T; ++t) {
      std::copy(kc[t].begin(), kc[t].end(), kCell.begin() + o);
      std::copy(kw[t].begin(), kw[t].end(), kW.begin() + o);
      o += kc[t].size();
    }
  }
  // ---- per-face volume weights (UG/Face/Face.cpp:54-58), face centroids, contact angle per boundary face
  std::vector<double> vw(F, 1.), fc(2 * (size_t)F), bfTheta;
  for (int f = 0; f < F; ++f) {
    if (m->fR[f] >= 0) vw[f] = m->vol[m->fR[f]] / (m->vol[m->fR[f]] + m->vol[m->fL[f]]);
    fc[f] = m->fCx[f]; fc[(size_t)F + f] = m->fCy[f];
    if (m->fR[f] < 0) {
      const int p = m->fPatch[f];
      bfTheta.push_back(p >= 0 && p < (int)mp->patchTheta.size() ? mp->patchTheta[p] : M_PI_2);
    }
  }
  cudaStream_t st = c->stream;
  PHB_CHECK(mp->stPtr.upload(stPtr, st)); PHB_CHECK(mp->stIdx.upload(stIdx, st));
  PHB_CHECK(mp->stKappa.upload(cK, st)); PHB_CHECK(mp->stGrad.upload(cG, st));
  PHB_CHECK(mp->kPtr.upload(kCnt, st)); PHB_CHECK(mp->kCell.upload(kCell, st)); PHB_CHECK(mp->kW.upload(kW, st));
  PHB_CHECK(mp->faceVW.upload(vw, st)); PHB_CHECK(mp->faceC.upload(fc, st)); PHB_CHECK(mp->bfTheta.upload(bfTheta, st));
  PHB_CUDA(cudaStreamSynchronize(st));
  mp->built = true;
  return PHB_OK;
}

int face_to_cell_weighted(phb_multiphase *mp, phb_field *v, const double *cellW, const double *faceW) {
  phb_mesh *m = mp->m;
  const View M = view(m);
  const int grid = (int)std::max<long long>(1, std::min<long long>(((long long)m->sell.nSlices * 32 + kThreads - 1) / kThreads,
                                                                    (long long)m->ctx->numSMs * 8));
  PHB_LAUNCH(m->ctx, k_face_to_cell_weighted, grid, kThreads, 0, M, v->faces.p, cellW, faceW, v->cells.p);
  if (m->nBCells)
    PHB_LAUNCH(m->ctx, k_face_to_cell_weighted_bnd, (m->nBCells + 255) / 256, 256, 0, M, v->faces.p, cellW, faceW, v->cells.p);
  return PHB_OK;
}

// ScalarGradient::computeFaces (UF/ScalarGradient.cpp:34-47)
__global__ void k_grad_faces_only(View M, const double *__restrict__ phiC, const double *__restrict__ phiF, double *__restrict__ gF) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < M.nFaces; f += gridDim.x * blockDim.x) {
    const int l = M.fL[f], r = M.fR[f];
    const double d = (r >= 0 ? phiC[r] : phiF[f]) - phiC[l];
    gF[f] = d * M.fQx[f];
    gF[(size_t)M.nFaces + f] = d * M.fQy[f];
  }
}

// FractionalStepMultiphase::updateProperties (:175-218)
int update_properties(phb_multiphase *mp) {
  phb_mesh *m = mp->m;
  phb_ctx *c = m->ctx;
  const View M = view(m);
  const int N = m->nLocal, F = m->nFaces, nDev = m->nDev;
  PHB_CHECK(phb_field_save_previous(mp->rho));
  PHB_CHECK(phb_field_save_previous(mp->mu));    // the reference saves mu after sg; both read nothing in between
  PHB_LAUNCH(c, k_properties, grid_flat(c, N), kThreads, 0, (long long)N, mp->gamma->cells.p, mp->rho1, mp->rho2, mp->mu1,
             mp->mu2, mp->rho->cells.p, mp->mu->cells.p);
  PHB_LAUNCH(c, k_properties, grid_flat(c, F), kThreads, 0, (long long)F, mp->gamma->faces.p, mp->rho1, mp->rho2, mp->mu1,
             mp->mu2, mp->rho->faces.p, mp->mu->faces.p);
  // gravity source
  PHB_LAUNCH(c, k_grad_faces_only, grid_flat(c, F), kThreads, 0, M, mp->rho->cells.p, mp->rho->faces.p, mp->gradRho->faces.p);
  PHB_LAUNCH(c, k_sg_faces, grid_flat(c, F), kThreads, 0, F, mp->faceC.p, mp->gx, mp->gy, mp->gradRho->faces.p, mp->sg->faces.p);
  PHB_CHECK(face_to_cell_weighted(mp, mp->sg, mp->rho->cells.p, mp->rho->faces.p));
  // surface tension: Celeste::computeFaceInterfaceForces (Celeste.cpp:10-20)
  PHB_LAUNCH(c, k_smooth, grid_flat(c, N), kThreads, 0, N, mp->kPtr.p, mp->kCell.p, mp->kW.p, mp->gamma->cells.p,
             mp->gammaTilde->cells.p);
  PHB_CHECK(phb::field_set_boundary_faces(mp->gammaTilde));
  PHB_LAUNCH(c, k_stencil_grad, grid_flat(c, N), kThreads, 0, N, nDev, mp->stPtr.p, mp->stIdx.p, mp->stGrad.p,
             mp->gammaTilde->cells.p, mp->gammaTilde->faces.p, mp->gradGammaTilde->cells.p);
  PHB_LAUNCH(c, k_normals, grid_flat(c, N), kThreads, 0, N, nDev, mp->eps * mp->eps, mp->gradGammaTilde->cells.p, mp->n->cells.p);
  if (m->nBFaces)
    PHB_LAUNCH(c, k_normals_bnd, (m->nBFaces + 255) / 256, 256, 0, M, mp->bfTheta.p, mp->n->cells.p, mp->n->faces.p);
  PHB_LAUNCH(c, k_curvature, grid_flat(c, N), kThreads, 0, N, nDev, F, mp->stPtr.p, mp->stIdx.p, mp->stKappa.p, mp->n->cells.p,
             mp->n->faces.p, mp->kappa->cells.p);
  PHB_LAUNCH(c, k_curvature_faces, grid_flat(c, F), kThreads, 0, M, mp->faceVW.p, mp->n->cells.p, mp->kappa->cells.p,
             mp->kappa->faces.p);
  PHB_LAUNCH(c, k_fst_faces, grid_flat(c, F), kThreads, 0, F, mp->sigma, mp->kappa->faces.p, mp->gradGamma->faces.p, mp->fst->faces.p);
  PHB_CHECK(face_to_cell_weighted(mp, mp->fst, mp->rho->cells.p, mp->rho->faces.p));
  return PHB_OK;
}

}  // namespace

extern "C" {

int phb_mp_create(phb_mesh *m, double rho1, double rho2, double mu1, double mu2, double sigma, double gx, double gy,
                  double kernelRadius, phb_multiphase **out) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(m && out && rho1 > 0. && rho2 > 0. && mu1 > 0. && mu2 > 0. && kernelRadius > 0., "phb_mp_create: bad argument");
  PHB_REQUIRE(m->finalized, "phb_mp_create: mesh is not finalized");
  if (m->ctx->nProcs > 1 || !m->identityCells) {
    set_error("phb_mp_create: the multiphase time step runs on one rank (n, kappa, sg, fst halos are not built)");
    return PHB_ERR_UNSUPPORTED;
  }
  phb_multiphase *mp = new phb_multiphase();
  mp->m = m; mp->rho1 = rho1; mp->rho2 = rho2; mp->mu1 = mu1; mp->mu2 = mu2; mp->sigma = sigma; mp->gx = gx; mp->gy = gy;
  mp->radius = kernelRadius;
  mp->patchTheta.assign(m->patchNames.size(), M_PI_2);
  struct { phb_field **f; int nc; const char *name; } fields[] = {
      {&mp->u, 2, "u"}, {&mp->p, 1, "p"}, {&mp->gradP, 2, "gradP"}, {&mp->gamma, 1, "gamma"}, {&mp->gradGamma, 2, "gradGamma"},
      {&mp->rho, 1, "rho"}, {&mp->mu, 1, "mu"}, {&mp->beta, 1, "beta"}, {&mp->sg, 2, "sg"}, {&mp->fst, 2, "fst"},
      {&mp->kappa, 1, "kappa"}, {&mp->gammaTilde, 1, "gammaTilde"}, {&mp->gradGammaTilde, 2, "gradGammaTilde"}, {&mp->n, 2, "n"},
      {&mp->gradRho, 2, "gradRho"}, {&mp->force, 2, "force"}, {&mp->dtRho, 1, "dtRho"}};
  for (auto &e : fields) PHB_CHECK(phb_field_create(m, e.nc, e.name, e.f));
  PHB_CHECK(phb_eqn_create(m, 1, &mp->gammaEqn)); PHB_CHECK(phb_eqn_create(m, 2, &mp->uEqn)); PHB_CHECK(phb_eqn_create(m, 1, &mp->pEqn));
  PHB_CHECK(phb_solver_create(m->ctx, &mp->gammaSolver)); PHB_CHECK(phb_solver_create(m->ctx, &mp->uSolver));
  PHB_CHECK(phb_solver_create(m->ctx, &mp->pSolver));
  PHB_CHECK(mp->out.alloc(4)); PHB_CHECK(mp->out.zero(m->ctx->stream));
  *out = mp;
  return PHB_OK;
  PHB_TRY_END
}

int phb_mp_destroy(phb_multiphase *mp) {
  if (!mp) return PHB_OK;
  for (phb_field *f : {mp->u, mp->p, mp->gradP, mp->gamma, mp->gradGamma, mp->rho, mp->mu, mp->beta, mp->sg, mp->fst, mp->kappa,
                       mp->gammaTilde, mp->gradGammaTilde, mp->n, mp->gradRho, mp->force, mp->dtRho}) phb_field_destroy(f);
  phb_eqn_destroy(mp->gammaEqn); phb_eqn_destroy(mp->uEqn); phb_eqn_destroy(mp->pEqn);
  phb_solver_destroy(mp->gammaSolver); phb_solver_destroy(mp->uSolver); phb_solver_destroy(mp->pSolver);
  delete mp;
  return PHB_OK;
}

phb_field *phb_mp_field(phb_multiphase *mp, const char *name) {
  if (!mp || !name) return nullptr;
  for (phb_field *f : {mp->u, mp->p, mp->gradP, mp->gamma, mp->gradGamma, mp->rho, mp->mu, mp->beta, mp->sg, mp->fst, mp->kappa,
                       mp->gammaTilde, mp->gradGammaTilde, mp->n, mp->gradRho})
    if (f->name == name) return f;
  return nullptr;
}
phb_eqn *phb_mp_eqn(phb_multiphase *mp, const char *name) {
  if (!mp || !name) return nullptr;
  return !strcmp(name, "gammaEqn") ? mp->gammaEqn : !strcmp(name, "uEqn") ? mp->uEqn : !strcmp(name, "pEqn") ? mp->pEqn : nullptr;
}
phb_solver *phb_mp_solver(phb_multiphase *mp, const char *name) {
  if (!mp || !name) return nullptr;
  return !strcmp(name, "gammaEqn") ? mp->gammaSolver : !strcmp(name, "uEqn") ? mp->uSolver : !strcmp(name, "pEqn") ? mp->pSolver : nullptr;
}

// keys: "eps" (Solver.eps), "kernelType" (0 peskin, 1 pow6, 2 pow8), "warmStart", "contactAngle:<patch>" (degrees)
int phb_mp_setup(phb_multiphase *mp, const char *key, double value) {
  PHB_REQUIRE(mp && key, "phb_mp_setup: NULL argument");
  PHB_REQUIRE(!mp->built, "phb_mp_setup: call before phb_mp_initialize");
  if (!strcmp(key, "eps")) mp->eps = value;
  else if (!strcmp(key, "kernelType")) mp->kernelType = (int)value;
  else if (!strcmp(key, "warmStart")) mp->warmStart = value != 0.;
  else if (!strncmp(key, "contactAngle:", 13)) {
    const int p = phb_mesh_patch_id(mp->m, key + 13);
    PHB_REQUIRE(p >= 0, "phb_mp_setup: no patch \"%s\"", key + 13);
    mp->patchTheta[p] = value * M_PI / 180.;
  } else PHB_REQUIRE(false, "phb_mp_setup: unknown key \"%s\"", key);
  return PHB_OK;
}

// FractionalStepMultiphase::initialize (:52-58)
int phb_mp_initialize(phb_multiphase *mp) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(mp, "phb_mp_initialize: NULL argument");
  if (!mp->built) PHB_CHECK(build_celeste(mp));
  PHB_CHECK(phb::field_interpolate_faces(mp->u));
  PHB_CHECK(phb::field_set_boundary_faces(mp->p));
  bool neumann = false;
  PHB_CHECK(phb::field_all_neumann(mp->p, &neumann));
  PHB_CHECK(phb_solver_setup(mp->pSolver, "nullSpace", neumann ? "constant" : "none"));
  PHB_CHECK(phb::field_gradient(mp->gamma, mp->gradGamma));
  PHB_CHECK(update_properties(mp));
  return launch_status(mp->m->ctx);
  PHB_TRY_END
}

// stats: [itersGamma, itersU, itersP, relresGamma, relresU, relresP, maxDivergence, maxCourant]
int phb_mp_step(phb_multiphase *mp, double dt, double stats[8]) {
  PHB_TRY_BEGIN
  PHB_REQUIRE(mp && dt > 0. && mp->built, "phb_mp_step: initialize first, dt > 0");
  phb_mesh *m = mp->m;
  phb_ctx *c = m->ctx;
  const View M = view(m);
  const int N = m->nLocal, F = m->nFaces, nDev = m->nDev;
  int itG = 0, itU = 0, itP = 0;
  double rrG = 0., rrU = 0., rrP = 0.;
  // ---- solveGammaEqn (:78-104)
  PHB_CHECK(phb_cicsam_weights(mp->u, mp->gamma, mp->gradGamma, dt, mp->beta));
  PHB_CHECK(phb_field_save_previous(mp->gamma));
  PHB_CHECK(phb_eqn_zero(mp->gammaEqn));
  PHB_CHECK(phb_assemble_ddt(mp->gammaEqn, mp->gamma, 1., nullptr, dt, +1.));
  PHB_CHECK(phb_assemble_cicsam_div(mp->gammaEqn, mp->u, mp->gamma, mp->beta, 0.5, +1.));
  PHB_CHECK(phb_eqn_solve(mp->gammaEqn, mp->gammaSolver, mp->gamma, mp->warmStart, &itG, &rrG));
  PHB_CHECK(phb::field_interpolate_faces(mp->gamma));
  PHB_CHECK(phb::field_gradient(mp->gamma, mp->gradGamma));
  // ---- updateProperties (:175-218)
  PHB_CHECK(update_properties(mp));
  // ---- solveUEqn (:106-150)
  PHB_CHECK(phb_field_save_previous(mp->u));
  PHB_CHECK(face_to_cell_weighted(mp, mp->gradP, mp->rho->cells.p, mp->rho->faces0.p));
  PHB_LAUNCH(c, k_force, grid_flat(c, 2 * N), kThreads, 0, N, nDev, mp->fst->cells.p, mp->sg->cells.p, mp->gradP->cells.p,
             mp->force->cells.p);
  PHB_CHECK(phb_eqn_zero(mp->uEqn));
  PHB_CHECK(phb_assemble_ddt(mp->uEqn, mp->u, 1., nullptr, dt, +1.));
  PHB_CHECK(phb_assemble_dive(mp->uEqn, mp->u, mp->u, 0.5, +1.));
  PHB_CHECK(phb_eqn_scale_rows(mp->uEqn, mp->rho));
  PHB_CHECK(phb_assemble_laplacian(mp->uEqn, 0., mp->mu, mp->u, 0.5, -1.));
  PHB_CHECK(phb_assemble_src(mp->uEqn, mp->force, -1.));
  PHB_CHECK(phb_eqn_solve(mp->uEqn, mp->uSolver, mp->u, mp->warmStart, &itU, &rrU));
  PHB_LAUNCH(c, k_axpy_over_rho, grid_flat(c, 2 * N), kThreads, 0, N, nDev, dt, mp->rho->cells.p, mp->gradP->cells.p, mp->u->cells.p);
  PHB_LAUNCH(c, k_sum_fields, grid_flat(c, 2 * (long long)nDev), kThreads, 0, 2 * (long long)nDev, mp->fst->cells.p, mp->sg->cells.p,
             mp->force->cells.p);
  PHB_LAUNCH(c, k_sum_fields, grid_flat(c, 2 * (long long)F), kThreads, 0, 2 * (long long)F, mp->fst->faces.p, mp->sg->faces.p,
             mp->force->faces.p);
  PHB_LAUNCH(c, k_momentum_faces, grid_flat(c, F), kThreads, 0, M, mp->faceVW.p, dt, mp->u->cells.p, mp->rho->cells.p,
             mp->rho->faces.p, mp->force->cells.p, mp->force->faces.p, mp->u->faces.p);
  PHB_CHECK(phb::field_face_types(mp->u));
  if (m->nBFaces)
    PHB_LAUNCH(c, k_momentum_faces_bnd, (m->nBFaces + 255) / 256, 256, 0, M, mp->u->dBfType.p, dt, mp->u->cells.p,
               mp->rho->cells.p, mp->rho->faces.p, mp->force->cells.p, mp->force->faces.p, mp->u->faces.p);
  // ---- solvePEqn (:152-163)
  PHB_LAUNCH(c, k_recip, grid_flat(c, nDev), kThreads, 0, (long long)nDev, dt, mp->rho->cells.p, mp->dtRho->cells.p);
  PHB_LAUNCH(c, k_recip, grid_flat(c, F), kThreads, 0, (long long)F, dt, mp->rho->faces.p, mp->dtRho->faces.p);
  PHB_CHECK(phb_eqn_zero(mp->pEqn));
  PHB_CHECK(phb_assemble_laplacian(mp->pEqn, 0., mp->dtRho, mp->p, -1., +1.));
  PHB_CHECK(phb_assemble_src_div(mp->pEqn, mp->u, -1.));
  PHB_CHECK(phb_eqn_solve(mp->pEqn, mp->pSolver, mp->p, mp->warmStart, &itP, &rrP));
  PHB_CHECK(phb::field_set_boundary_faces(mp->p));
  PHB_LAUNCH(c, k_grad_faces_only, grid_flat(c, F), kThreads, 0, M, mp->p->cells.p, mp->p->faces.p, mp->gradP->faces.p);
  PHB_CHECK(face_to_cell_weighted(mp, mp->gradP, mp->rho->cells.p, mp->rho->faces.p));
  // ---- correctVelocity (:165-173)
  PHB_LAUNCH(c, k_axpy_over_rho, grid_flat(c, 2 * N), kThreads, 0, N, nDev, -dt, mp->rho->cells.p, mp->gradP->cells.p, mp->u->cells.p);
  PHB_LAUNCH(c, k_axpy_over_rho_faces, grid_flat(c, 2 * F), kThreads, 0, F, -dt, mp->rho->faces.p, mp->gradP->faces.p, mp->u->faces.p);
  // ---- diagnostics (:69-71)
  PHB_CHECK(phb::field_flux_max(mp->u, 0, dt, mp->scratch, mp->partials, mp->ticket, mp->out.p));
  PHB_CHECK(phb::field_flux_max(mp->u, 1, dt, mp->scratch, mp->partials, mp->ticket, mp->out.p + 1));
  PHB_CUDA(cudaMemcpyAsync(c->pinned + 64, mp->out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  if (stats) {
    stats[0] = itG; stats[1] = itU; stats[2] = itP; stats[3] = rrG; stats[4] = rrU; stats[5] = rrP;
    stats[6] = c->pinned[64]; stats[7] = c->pinned[65];
  }
  return launch_status(c);
  PHB_TRY_END
}

}  // extern "C"
