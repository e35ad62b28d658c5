// piso.cu -- device-resident PISO time step ("phasePiso" of the north star).
//
// The mounted snapshot no longer ships a PISO/SIMPLE module;
// what remains of it is README.md:26-37 (the two equations), the commented body
// of relax() (UE/ScalarFiniteVolumeEquation.cpp:45-55) and the legacy case keys
// numInnerIterations / numPressureCorrections / momentumRelaxation /
// pressureCorrectionRelaxation (Examples/LidDrivenCavity/case/case.info:12-15).
// This driver rebuilds it from those pieces with the same fv:: kernels; its parity
// is SELF-CONSISTENCY (mass conservation, agreement with the fractional-step
// steady state), not a reference comparison -- there is nothing to compare with.
//
//   per time step:  u.savePreviousTimeStep
//   per inner iteration:
//     uEqn  = (fv::ddt(rho,u,dt) + fv::div(rho*u,u) == fv::laplacian(mu,u) - fv::grad(p)); relax(w_u); solve
//     d     = V / a_P                                       (momentum diagonal)
//     per pressure correction:
//       u_f   = interp(u) - d_f [(grad p)_f - interp(grad p)]      (Rhie-Chow)
//       m     = sum_f rho u_f . S_f
//       pCorrEqn = (fv::laplacian(rho*d, pCorr) == m); solve
//       p    += w_p pCorr ;  u -= d grad(pCorr) (cells and faces) ;  grad p recomputed
#include <cmath>

#include "comm.cuh"
#include "fv.cuh"
#include "kernels.cuh"
#include "solver.cuh"

namespace {
constexpr int kThreads = 256;

// d_P = V_P / a_P from the diagonal (entry 0 of every row) of the assembled momentum equation
__global__ void k_diag_to_d(int nRows, const int *__restrict__ sliceOff, const double *__restrict__ vals,
                            const double *__restrict__ vol, double *__restrict__ d) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nRows) return;
  d[row] = vol[row] / vals[(size_t)sliceOff[row >> 5] + (row & 31)];
}
// y[c][i] += a * w[i] * x[c][i]
__global__ void k_axpy_weighted(long long n, int nc, long long ldy, long long ldx, double a,
                                const double *__restrict__ w, const double *__restrict__ x, double *__restrict__ y) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n * nc;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / n, j = i - c * n;
    y[c * ldy + j] += a * w[j] * x[c * ldx + j];
  }
}
// Rhie-Chow correction on interior faces: u_f -= d_f [ (grad p)_f - (w gradP_l + (1-w) gradP_r) ]
__global__ void k_rhie_chow(int nIF, const int *__restrict__ ifFace, const int *__restrict__ fL,
                            const int *__restrict__ fR, const double *__restrict__ fW, int nDev, int nFaces,
                            const double *__restrict__ dF, const double *__restrict__ gC,
                            const double *__restrict__ gF, double *__restrict__ uF) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nIF; i += gridDim.x * blockDim.x) {
    const int f = ifFace[i], l = fL[f], r = fR[f];
    const double w = fW[f], df = dF[f];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const double gbar = w * gC[(size_t)c * nDev + l] + (1. - w) * gC[(size_t)c * nDev + r];
      uF[(size_t)c * nFaces + f] -= df * (gF[(size_t)c * nFaces + f] - gbar);
    }
  }
}
int grid_for(const phb_ctx *c, long long n) {
  return (int)std::max<long long>(1, std::min<long long>((n + kThreads - 1) / kThreads, (long long)c->numSMs * 8));
}
}  // namespace

struct phb_piso {
  phb_mesh *m = nullptr;
  double rho = 1., mu = 1.;
  int numInner = 1, numCorr = 1;
  double omegaU = 0.8, omegaP = 0.2;
  phb_field *u = nullptr, *p = nullptr, *pCorr = nullptr, *gradP = nullptr, *gradPCorr = nullptr, *rhoU = nullptr,
            *d = nullptr, *rhoD = nullptr;
  phb_eqn *uEqn = nullptr, *pCorrEqn = nullptr;
  phb_solver *uSolver = nullptr, *pSolver = nullptr;
  phb::DevBuf<double> scratch, partials, out;
  phb::DevBuf<unsigned> ticket;
};

extern "C" {

int phb_piso_create(phb_mesh *m, double rho, double mu, phb_piso **out) {
  PHB_REQUIRE(m && out && rho > 0., "phb_piso_create: bad argument");
  PHB_REQUIRE(m->finalized, "phb_piso_create: mesh is not finalized");
  phb_piso *s = new phb_piso();
  s->m = m; s->rho = rho; s->mu = mu;
  PHB_CHECK(phb_field_create(m, 2, "u", &s->u)); PHB_CHECK(phb_field_create(m, 1, "p", &s->p));
  PHB_CHECK(phb_field_create(m, 1, "pCorr", &s->pCorr)); PHB_CHECK(phb_field_create(m, 2, "gradP", &s->gradP));
  PHB_CHECK(phb_field_create(m, 2, "gradPCorr", &s->gradPCorr)); PHB_CHECK(phb_field_create(m, 2, "rhoU", &s->rhoU));
  PHB_CHECK(phb_field_create(m, 1, "d", &s->d)); PHB_CHECK(phb_field_create(m, 1, "rhoD", &s->rhoD));
  PHB_CHECK(phb_eqn_create(m, 2, &s->uEqn)); PHB_CHECK(phb_eqn_create(m, 1, &s->pCorrEqn));
  PHB_CHECK(phb_solver_create(m->ctx, &s->uSolver)); PHB_CHECK(phb_solver_create(m->ctx, &s->pSolver));
  PHB_CHECK(s->out.alloc(4)); PHB_CHECK(s->out.zero(m->ctx->stream));
  *out = s;
  return PHB_OK;
}
int phb_piso_destroy(phb_piso *s) {
  if (!s) return PHB_OK;
  for (phb_field *f : {s->u, s->p, s->pCorr, s->gradP, s->gradPCorr, s->rhoU, s->d, s->rhoD}) phb_field_destroy(f);
  phb_eqn_destroy(s->uEqn); phb_eqn_destroy(s->pCorrEqn);
  phb_solver_destroy(s->uSolver); phb_solver_destroy(s->pSolver);
  delete s;
  return PHB_OK;
}
phb_field *phb_piso_field(phb_piso *s, const char *name) {
  if (!s || !name) return nullptr;
  if (!strcmp(name, "u")) return s->u;
  if (!strcmp(name, "p")) return s->p;
  if (!strcmp(name, "pCorr")) return s->pCorr;
  if (!strcmp(name, "gradP")) return s->gradP;
  if (!strcmp(name, "d")) return s->d;
  return nullptr;
}
phb_solver *phb_piso_solver(phb_piso *s, const char *name) {
  if (!s || !name) return nullptr;
  if (!strcmp(name, "uEqn")) return s->uSolver;
  if (!strcmp(name, "pCorrEqn")) return s->pSolver;
  return nullptr;
}
// keys of the legacy case file: numInnerIterations numPressureCorrections momentumRelaxation pressureCorrectionRelaxation
int phb_piso_setup(phb_piso *s, const char *key, double value) {
  PHB_REQUIRE(s && key, "phb_piso_setup: NULL argument");
  if (!strcmp(key, "numInnerIterations")) s->numInner = (int)value;
  else if (!strcmp(key, "numPressureCorrections")) s->numCorr = (int)value;
  else if (!strcmp(key, "momentumRelaxation")) s->omegaU = value;
  else if (!strcmp(key, "pressureCorrectionRelaxation")) s->omegaP = value;
  else PHB_REQUIRE(false, "phb_piso_setup: unknown key \"%s\"", key);
  PHB_REQUIRE(s->numInner >= 1 && s->numCorr >= 1 && s->omegaU > 0. && s->omegaP > 0., "phb_piso_setup: bad value");
  return PHB_OK;
}

// pCorr takes the boundary TYPES of p with zero reference values; d and rho*d are zero-gradient
int phb_piso_initialize(phb_piso *s) {
  PHB_REQUIRE(s, "phb_piso_initialize: NULL argument");
  for (size_t i = 0; i < s->p->bc.size(); ++i) {
    s->pCorr->bc[i].type = s->p->bc[i].type;
    s->pCorr->bc[i].vx = 0.;
  }
  s->pCorr->bcDirty = true;
  PHB_CHECK(phb::field_send_messages(s->u));
  PHB_CHECK(phb::field_interpolate_faces(s->u));
  PHB_CHECK(phb::field_set_boundary_faces(s->p));
  PHB_CHECK(phb::field_gradient(s->p, s->gradP));
  bool neumann = false;
  PHB_CHECK(phb::field_all_neumann(s->p, &neumann));
  PHB_CHECK(phb_solver_setup(s->pSolver, "nullSpace", neumann ? "constant" : "none"));
  return PHB_OK;
}

// stats: [itersU (last), itersPCorr (sum), relresU, relresP, maxMassImbalance, maxCourant]
int phb_piso_step(phb_piso *s, double dt, double stats[6]) {
  PHB_REQUIRE(s && dt > 0., "phb_piso_step: bad argument");
  phb_mesh *m = s->m;
  phb_ctx *c = m->ctx;
  int itU = 0, itP = 0, it = 0;
  double rrU = 0., rrP = 0.;
  PHB_CHECK(phb_field_save_previous(s->u));
  for (int inner = 0; inner < s->numInner; ++inner) {
    // ---- momentum predictor
    PHB_CHECK(phb_field_fill(s->rhoU, 0., 0.));
    PHB_CHECK(phb::field_axpy_faces(s->rhoU, s->rho, s->u));
    s->rhoU->hasOld = true;  // theta = 1: the old flux is never read
    PHB_CHECK(phb_eqn_zero(s->uEqn));
    PHB_CHECK(phb_assemble_ddt(s->uEqn, s->u, s->rho, nullptr, dt, +1.));
    PHB_CHECK(phb_assemble_div(s->uEqn, s->rhoU, s->u, 1., +1.));
    PHB_CHECK(phb_assemble_laplacian(s->uEqn, s->mu, nullptr, s->u, -1., -1.));
    PHB_CHECK(phb_assemble_src(s->uEqn, s->gradP, +1.));  // == (... - fv::grad(p))
    PHB_CHECK(phb_eqn_relax(s->uEqn, s->u, s->omegaU));
    PHB_LAUNCH(c, k_diag_to_d, (m->nLocal + 255) / 256, 256, 0, m->nLocal, m->sell.sliceOff.p, s->uEqn->vals.p,
               m->dVol.p, s->d->cells.p);
    PHB_CHECK(phb_eqn_solve(s->uEqn, s->uSolver, s->u, 1, &itU, &rrU));
    PHB_CHECK(phb::field_send_messages(s->u));
    PHB_CHECK(phb::field_send_messages(s->d));
    PHB_CHECK(phb::field_interpolate_faces(s->d));
    // rho*d on cells and faces
    PHB_CHECK(phb_field_fill(s->rhoD, 0., 0.));
    PHB_CHECK(phb::field_axpy_cells(s->rhoD, s->rho, s->d));
    PHB_CHECK(phb::field_axpy_faces(s->rhoD, s->rho, s->d));
    for (int corr = 0; corr < s->numCorr; ++corr) {
      // ---- Rhie-Chow face velocity and mass imbalance
      PHB_CHECK(phb::field_interpolate_faces(s->u));
      if (m->nIFaces)
        PHB_LAUNCH(c, k_rhie_chow, grid_for(c, m->nIFaces), kThreads, 0, m->nIFaces, m->dIfFace.p, m->dFL.p, m->dFR.p,
                   m->dFW.p, m->nDev, m->nFaces, s->d->faces.p, s->gradP->cells.p, s->gradP->faces.p, s->u->faces.p);
      // ---- pCorrEqn = (fv::laplacian(rho*d, pCorr) == m)
      PHB_CHECK(phb_field_fill(s->pCorr, 0., 0.));
      PHB_CHECK(phb_eqn_zero(s->pCorrEqn));
      PHB_CHECK(phb_assemble_laplacian(s->pCorrEqn, 0., s->rhoD, s->pCorr, -1., +1.));
      PHB_CHECK(phb_field_fill(s->rhoU, 0., 0.));
      PHB_CHECK(phb::field_axpy_faces(s->rhoU, s->rho, s->u));
      PHB_CHECK(phb_assemble_src_div(s->pCorrEqn, s->rhoU, -1.));
      PHB_CHECK(phb_eqn_solve(s->pCorrEqn, s->pSolver, s->pCorr, 0, &it, &rrP));
      itP += it;
      PHB_CHECK(phb::field_send_messages(s->pCorr));
      PHB_CHECK(phb::field_set_boundary_faces(s->pCorr));
      PHB_CHECK(phb::field_gradient(s->pCorr, s->gradPCorr));
      // ---- corrections: p += w_p pCorr ; u -= d grad(pCorr) on cells and faces
      PHB_CHECK(phb::field_axpy_cells(s->p, s->omegaP, s->pCorr));
      PHB_LAUNCH(c, k_axpy_weighted, grid_for(c, 2LL * m->nLocal), kThreads, 0, (long long)m->nLocal, 2,
                 (long long)m->nDev, (long long)m->nDev, -1., s->d->cells.p, s->gradPCorr->cells.p, s->u->cells.p);
      PHB_LAUNCH(c, k_axpy_weighted, grid_for(c, 2LL * m->nFaces), kThreads, 0, (long long)m->nFaces, 2,
                 (long long)m->nFaces, (long long)m->nFaces, -1., s->d->faces.p, s->gradPCorr->faces.p, s->u->faces.p);
      PHB_CHECK(phb::field_send_messages(s->p));
      PHB_CHECK(phb::field_send_messages(s->u));
      PHB_CHECK(phb::field_set_boundary_faces(s->p));
      PHB_CHECK(phb::field_gradient(s->p, s->gradP));
    }
  }
  PHB_CHECK(phb::field_flux_max(s->u, 0, dt, s->scratch, s->partials, s->ticket, s->out.p));
  PHB_CHECK(phb::field_flux_max(s->u, 1, dt, s->scratch, s->partials, s->ticket, s->out.p + 1));
  PHB_CUDA(cudaMemcpyAsync(c->pinned + 64, s->out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  if (stats) {
    stats[0] = itU; stats[1] = itP; stats[2] = rrU; stats[3] = rrP;
    stats[4] = c->pinned[64]; stats[5] = c->pinned[65];
  }
  return PHB_OK;
}

}  // extern "C"
