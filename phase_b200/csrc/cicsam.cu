// cicsam.cu -- CICSAM interface-capturing advection (A9): face weights beta_f, the VOF
// advection operator and the density-weighted momentum flux.
// Reference: src/2D/Unstructured/FiniteVolume/Discretization/Cicsam.cpp
//   hc / uq                        :5-17     Hyper-C and ULTIMATE-QUICKEST normalised-variable limiters
//   faceInterpolationWeights       :19-66    donor/acceptor by flux sign, blend by interface angle
//   computeMomentumFlux            :69-87
//   div(u, gamma, beta, theta)     :89-138   A_P,donor += theta (1-b) F ; A_P,acceptor += theta b F
// Two launches for the weights: a cell-parallel Courant pass (the donor's Courant number is a
// per-cell quantity the reference recomputes for every face), then a face-parallel pass.
#include <cmath>

#include "fv.cuh"
#include "kernels.cuh"

namespace {
constexpr int kThreads = 256;

__device__ __forceinline__ double clampd(double v, double lo, double hi) { return fmax(fmin(v, hi), lo); }
__device__ __forceinline__ double hc(double g, double co) { return g >= 0. && g <= 1. ? fmin(1., g / co) : g; }
__device__ __forceinline__ double uq(double g, double co) {
  return g >= 0. && g <= 1. ? fmin((8. * co * g + (1. - co) * (6. * g + 3.)) / 8., hc(g, co)) : g;
}

struct View {
  const int *sliceOff, *col, *linkFace;
  int nRows, nSlices, nDev, nFaces, nBCells;
  const double *vol, *fSx, *fSy;
  const int *fL, *fR, *bcCell, *bcPtr, *bcFace;
};
View view(const phb_mesh *m) {
  View v;
  v.sliceOff = m->sell.sliceOff.p; v.col = m->sell.col.p; v.linkFace = m->dLinkFace.p;
  v.nRows = m->sell.nRows; v.nSlices = m->sell.nSlices; v.nDev = m->nDev; v.nFaces = m->nFaces; v.nBCells = m->nBCells;
  v.vol = m->dVol.p; v.fSx = m->dFSx.p; v.fSy = m->dFSy.p; v.fL = m->dFL.p; v.fR = m->dFR.p;
  v.bcCell = m->dBcCell.p; v.bcPtr = m->dBcPtr.p; v.bcFace = m->dBcFace.p;
  return v;
}

// co_P = sum over links of max(F / V dt, 0)
__global__ void k_courant(View M, const double *__restrict__ uF, double dt, double *__restrict__ co) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int slice = blockIdx.x * wpb + (threadIdx.x >> 5); slice < M.nSlices; slice += gridDim.x * wpb) {
    const int off = M.sliceOff[slice], wdt = (M.sliceOff[slice + 1] - off) >> 5, row = slice * 32 + lane;
    if (row >= M.nRows) continue;
    const double V = M.vol[row];
    double c = 0.;
    for (int k = 1; k < wdt; ++k) {
      const int lf = M.linkFace[(size_t)off + (size_t)k * 32 + lane];
      if (lf < 0) continue;
      const int f = lf >> 1;
      const double flux = ((lf & 1) ? -1. : 1.) * (uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f]);
      c += fmax(flux / V * dt, 0.);
    }
    co[row] = c;
  }
}
__global__ void k_courant_bnd(View M, const double *__restrict__ uF, double dt, double *__restrict__ co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  const double V = M.vol[row];
  double c = co[row];
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j];
    c += fmax((uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f]) / V * dt, 0.);
  }
  co[row] = c;
}

// beta_f on interior faces (0 on boundary faces)
__global__ void k_cicsam_weights(View M, const double *__restrict__ uF, const double *__restrict__ gam,
                                 const double *__restrict__ gradG, const double *__restrict__ co,
                                 const double *__restrict__ cx, const double *__restrict__ cy,
                                 double *__restrict__ beta) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < M.nFaces; f += gridDim.x * blockDim.x) {
    const int l = M.fL[f], r = M.fR[f];
    if (r < 0) { beta[f] = 0.; continue; }
    const double flux = uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f];
    const int d = flux > 0. ? l : r, a = flux <= 0. ? l : r;
    const double rcx = cx[a] - cx[d], rcy = cy[a] - cy[d];
    const double gx = gradG[d], gy = gradG[(size_t)M.nDev + d];
    const double gD = clampd(gam[d], 0., 1.), gA = clampd(gam[a], 0., 1.);
    const double gU = clampd(gA - 2. * (rcx * gx + rcy * gy), 0., 1.);
    double gDT = (gD - gU) / (gA - gU);
    if (!isfinite(gDT)) gDT = 0.;
    const double coD = co[d];
    const double gm = sqrt(gx * gx + gy * gy), rm = sqrt(rcx * rcx + rcy * rcy);
    const double thetaF = acos(fabs((gx / gm) * (rcx / rm) + (gy / gm) * (rcy / rm)));
    const double psiF = fmin((cos(2. * thetaF) + 1.) / 2., 1.);
    const double gFT = psiF * hc(gDT, coD) + (1. - psiF) * uq(gDT, coD);
    double b = (gFT - gDT) / (1. - gDT);
    b = isfinite(b) ? fmax(fmin(1., b), 0.) : 0.;
    beta[f] = b;
  }
}

// cicsam::div accumulated into (vals, rhs) with sign
__global__ void k_cicsam_div(View M, double *__restrict__ vals, double *__restrict__ rhs,
                             const double *__restrict__ uF, const double *__restrict__ beta,
                             const double *__restrict__ g0, double theta, double sign) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int slice = blockIdx.x * wpb + (threadIdx.x >> 5); slice < M.nSlices; slice += gridDim.x * wpb) {
    const int off = M.sliceOff[slice], wdt = (M.sliceOff[slice + 1] - off) >> 5, row = slice * 32 + lane;
    if (row >= M.nRows) continue;
    const size_t slot0 = (size_t)off + lane;
    double diag = 0., r = 0.;
    for (int k = 1; k < wdt; ++k) {
      const size_t slot = slot0 + (size_t)k * 32;
      const int lf = M.linkFace[slot];
      if (lf < 0) continue;
      const int f = lf >> 1, nb = M.col[slot];
      const double flux = ((lf & 1) ? -1. : 1.) * (uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f]);
      const double b = beta[f];
      const bool selfDonor = flux > 0.;
      const double cD = theta * (1. - b) * flux, cA = theta * b * flux;
      diag += selfDonor ? cD : cA;
      vals[slot] += sign * (selfDonor ? cA : cD);
      const double gF = (1. - b) * g0[selfDonor ? row : nb] + b * g0[selfDonor ? nb : row];
      r += (1. - theta) * flux * gF;
    }
    vals[slot0] += sign * diag;
    rhs[row] += sign * r;
  }
}
__global__ void k_cicsam_div_bnd(View M, const int *__restrict__ faceType, double *__restrict__ vals,
                                 double *__restrict__ rhs, const double *__restrict__ uF,
                                 const double *__restrict__ gF, const double *__restrict__ g0F,
                                 const double *__restrict__ g0, double theta, double sign) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M.nBCells) return;
  const int row = M.bcCell[i];
  const size_t slot0 = (size_t)M.sliceOff[row >> 5] + (row & 31);
  double diag = 0., r = 0.;
  for (int j = M.bcPtr[i]; j < M.bcPtr[i + 1]; ++j) {
    const int f = M.bcFace[j], t = faceType[f];
    const double flux = uF[f] * M.fSx[f] + uF[(size_t)M.nFaces + f] * M.fSy[f];
    if (t == PHB_FIXED) {
      r += theta * flux * gF[f];
      r += (1. - theta) * flux * g0F[f];
    } else if (t == PHB_NORMAL_GRADIENT) {
      diag += theta * flux;
      r += (1. - theta) * flux * g0[row];
    }
  }
  vals[slot0] += sign * diag;
  rhs[row] += sign * r;
}

__global__ void k_momentum_flux(View M, double rho1, double rho2, const double *__restrict__ uF,
                                const double *__restrict__ gam, const double *__restrict__ gamF,
                                const double *__restrict__ beta, double *__restrict__ out) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < M.nFaces; f += gridDim.x * blockDim.x) {
    const int l = M.fL[f], r = M.fR[f];
    const double ux = uF[f], uy = uF[(size_t)M.nFaces + f];
    double g;
    if (r >= 0) {
      const double flux = ux * M.fSx[f] + uy * M.fSy[f];
      const int d = flux > 0. ? l : r, a = flux <= 0. ? l : r;
      g = (1. - beta[f]) * gam[d] + beta[f] * gam[a];
    } else
      g = gamF[f];
    const double rho = rho1 + clampd(g, 0., 1.) * (rho2 - rho1);
    out[f] = rho * ux;
    out[(size_t)M.nFaces + f] = rho * uy;
  }
}

int grid_rows(const phb_ctx *c, const phb_mesh *m) {
  return (int)std::max<long long>(1, std::min<long long>(((long long)m->sell.nSlices * 32 + kThreads - 1) / kThreads,
                                                        (long long)c->numSMs * 8));
}
int grid_flat(const phb_ctx *c, long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + kThreads - 1) / kThreads, (long long)c->numSMs * 8));
}
}  // namespace

extern "C" {

int phb_cicsam_weights(const phb_field *u, const phb_field *gamma, const phb_field *gradGamma, double dt,
                       phb_field *beta) {
  PHB_REQUIRE(u && gamma && gradGamma && beta, "phb_cicsam_weights: NULL argument");
  phb_mesh *m = gamma->m;
  PHB_REQUIRE(u->m == m && gradGamma->m == m && beta->m == m, "phb_cicsam_weights: fields live on different meshes");
  PHB_REQUIRE(u->nComp == 2 && gamma->nComp == 1 && gradGamma->nComp == 2 && beta->nComp == 1 && dt > 0.,
              "phb_cicsam_weights: u, gradGamma vector; gamma, beta scalar; dt > 0");
  phb_ctx *c = m->ctx;
  const View M = view(m);
  if (!m->dCellC.p) {  // cell centroids in device numbering (uploaded on first use)
    std::vector<double> cc(2 * (size_t)m->nDev);
    for (int d = 0; d < m->nDev; ++d) { cc[d] = m->cCx[m->dev2cell[d]]; cc[(size_t)m->nDev + d] = m->cCy[m->dev2cell[d]]; }
    PHB_CHECK(m->dCellC.upload(cc, c->stream));
    PHB_CUDA(cudaStreamSynchronize(c->stream));
  }
  PHB_CHECK(m->dCo.alloc((size_t)m->nDev));
  PHB_CHECK(m->dCo.zero(c->stream));
  PHB_LAUNCH(c, k_courant, grid_rows(c, m), kThreads, 0, M, u->faces.p, dt, m->dCo.p);
  if (m->nBCells) PHB_LAUNCH(c, k_courant_bnd, (m->nBCells + 255) / 256, 256, 0, M, u->faces.p, dt, m->dCo.p);
  if (c->nProcs > 1) {  // the donor of a partition-boundary face may be a ghost: its Courant number comes from its owner
    phb_field tmp;      // view over dCo with the field halo machinery
    tmp.m = m; tmp.nComp = 1; tmp.cells.attach(m->dCo.p, (size_t)m->nDev);
    PHB_CHECK(phb::field_send_messages(&tmp));
  }
  PHB_LAUNCH(c, k_cicsam_weights, grid_flat(c, m->nFaces), kThreads, 0, M, u->faces.p, gamma->cells.p,
             gradGamma->cells.p, m->dCo.p, m->dCellC.p, m->dCellC.p + m->nDev, beta->faces.p);
  return PHB_OK;
}

int phb_assemble_cicsam_div(phb_eqn *e, const phb_field *u, const phb_field *cgamma, const phb_field *beta,
                            double theta, double sign) {
  PHB_REQUIRE(e && u && cgamma && beta, "phb_assemble_cicsam_div: NULL argument");
  phb_mesh *m = e->m;
  PHB_REQUIRE(u->m == m && cgamma->m == m && beta->m == m, "phb_assemble_cicsam_div: different meshes");
  PHB_REQUIRE(e->nComp == 1 && cgamma->nComp == 1 && u->nComp == 2 && beta->nComp == 1,
              "phb_assemble_cicsam_div: scalar equation/field, vector u");
  PHB_REQUIRE(cgamma->hasOld, "phb_assemble_cicsam_div: gamma needs a previous time step");
  phb_field *gamma = const_cast<phb_field *>(cgamma);
  PHB_CHECK(phb::field_face_types(gamma));
  phb_ctx *c = m->ctx;
  const View M = view(m);
  PHB_LAUNCH(c, k_cicsam_div, grid_rows(c, m), kThreads, 0, M, e->vals.p, e->rhs.p, u->faces.p, beta->faces.p,
             gamma->cells0.p, theta, sign);
  if (m->nBCells)
    PHB_LAUNCH(c, k_cicsam_div_bnd, (m->nBCells + 255) / 256, 256, 0, M, gamma->dFaceType.p, e->vals.p, e->rhs.p,
               u->faces.p, gamma->faces.p, gamma->faces0.p, gamma->cells0.p, theta, sign);
  return PHB_OK;
}

int phb_cicsam_momentum_flux(double rho1, double rho2, const phb_field *u, const phb_field *gamma,
                             const phb_field *beta, phb_field *rhoU) {
  PHB_REQUIRE(u && gamma && beta && rhoU, "phb_cicsam_momentum_flux: NULL argument");
  phb_mesh *m = gamma->m;
  PHB_REQUIRE(u->m == m && beta->m == m && rhoU->m == m && u->nComp == 2 && rhoU->nComp == 2 && gamma->nComp == 1,
              "phb_cicsam_momentum_flux: bad fields");
  PHB_LAUNCH(m->ctx, k_momentum_flux, grid_flat(m->ctx, m->nFaces), kThreads, 0, view(m), rho1, rho2, u->faces.p,
             gamma->cells.p, gamma->faces.p, beta->faces.p, rhoU->faces.p);
  return PHB_OK;
}

}  // extern "C"
