// fracstep.cu -- device-resident FractionalStep::solve: the caller of the hot
// path (US/FractionalStep.cpp:36-135), built from the fv:: kernels and the
// BiCGStab solver so that a whole time step runs without host<->device field
// traffic.  Fields, equations and solvers are exposed so callers can drive the
// pieces themselves (the C++ mirror in include/phase/ does).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "comm.cuh"
#include "fv.cuh"
#include "solver.cuh"

struct phb_fracstep {
  phb_mesh *m = nullptr;
  double rho = 1., mu = 1.;
  phb_field *u = nullptr, *p = nullptr, *gradP = nullptr;
  phb_eqn *uEqn = nullptr, *pEqn = nullptr;
  phb_solver *uSolver = nullptr, *pSolver = nullptr;
  phb::DevBuf<double> scratch, partials, out;
  phb::DevBuf<unsigned> ticket;
  bool warmStart = true;
  // initial guess of pEqn_: 0 = previous p (plain warm start), 1 = linear extrapolation 2 p^n - p^(n-1).  The same
  // extrapolation of the momentum predictors was measured and dropped: 7.55 instead of 7.15 uEqn_ iterations at 4M cells
  // (the predictor sequence carries the lagged pressure gradient and is not smooth in time); p gains 7.7 -> 7.35.
  int guessOrder = 1;
  phb::DevBuf<double> pPrev;  // p^(n-1), owned cells
  phb::DevBuf<double> uSave;  // cells of u while phb_fs_rebuild_faces borrows them
  int nStepsDone = 0;
  unsigned long long uTag = 0, pTag = 0;   // coefficient tags of the last one-pass assembly (0: assembled term by term)
  bool fusedAssembly = true;     // uEqn_ and pEqn_ each in one pass over the rows ("fusedAssembly" 0: one kernel per operator)
};

namespace {
// equal tags <=> equal coefficients of a one-pass assembly (never 0)
unsigned long long matrix_tag(int which, double dt, double gamma, unsigned bcVersion) {
  unsigned long long a, b;
  memcpy(&a, &dt, 8);
  memcpy(&b, &gamma, 8);
  unsigned long long h = 1469598103934665603ull;
  for (unsigned long long v : {(unsigned long long)which, a, b, (unsigned long long)bcVersion}) {
    h ^= v;
    h *= 1099511628211ull;
    h ^= h >> 29;
  }
  return h ? h : 1ull;
}
// x = 2 x - xPrev ; xPrev = old x        (owned cells of a scalar field)
__global__ void k_extrapolate(int n, double *__restrict__ x, double *__restrict__ xPrev, int apply) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double cur = x[i], old = xPrev[i];
    xPrev[i] = cur;
    if (apply) x[i] = 2. * cur - old;
  }
}
// y += a x on every face that is not a FIXED boundary face of the field (correctVelocity, US/FractionalStep.cpp:109-117)
__global__ void k_axpy_faces_nonfixed(long long nF, int nc, const int *__restrict__ faceType, double a,
                                      const double *__restrict__ x, double *__restrict__ y) {
  for (long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x; f < nF; f += (long long)gridDim.x * blockDim.x) {
    if (faceType[f] == PHB_FIXED) continue;
    for (int c = 0; c < nc; ++c) y[(size_t)c * nF + f] += a * x[(size_t)c * nF + f];
  }
}
}  // namespace

extern "C" {

int phb_fs_create(phb_mesh *m, double rho, double mu, phb_fracstep **out) {
  PHB_REQUIRE(m && out && rho > 0., "phb_fs_create: bad argument");
  PHB_REQUIRE(m->finalized, "phb_fs_create: mesh is not finalized");
  phb_fracstep *fs = new phb_fracstep();
  fs->m = m; fs->rho = rho; fs->mu = mu;
  PHB_CHECK(phb_field_create(m, 2, "u", &fs->u));
  PHB_CHECK(phb_field_create(m, 1, "p", &fs->p));
  PHB_CHECK(phb_field_create(m, 2, "gradP", &fs->gradP));
  PHB_CHECK(phb_eqn_create(m, 2, &fs->uEqn));
  PHB_CHECK(phb_eqn_create(m, 1, &fs->pEqn));
  PHB_CHECK(phb_solver_create(m->ctx, &fs->uSolver));
  PHB_CHECK(phb_solver_create(m->ctx, &fs->pSolver));
  PHB_CHECK(fs->out.alloc(4));
  PHB_CHECK(fs->out.zero(m->ctx->stream));
  *out = fs;
  return PHB_OK;
}

int phb_fs_destroy(phb_fracstep *fs) {
  if (!fs) return PHB_OK;
  phb_field_destroy(fs->u); phb_field_destroy(fs->p); phb_field_destroy(fs->gradP);
  phb_eqn_destroy(fs->uEqn); phb_eqn_destroy(fs->pEqn);
  phb_solver_destroy(fs->uSolver); phb_solver_destroy(fs->pSolver);
  delete fs;
  return PHB_OK;
}

phb_field *phb_fs_field(phb_fracstep *fs, const char *name) {
  if (!fs || !name) return nullptr;
  if (!strcmp(name, "u")) return fs->u;
  if (!strcmp(name, "p")) return fs->p;
  if (!strcmp(name, "gradP")) return fs->gradP;
  return nullptr;
}
phb_eqn *phb_fs_eqn(phb_fracstep *fs, const char *name) {
  if (!fs || !name) return nullptr;
  if (!strcmp(name, "uEqn")) return fs->uEqn;
  if (!strcmp(name, "pEqn")) return fs->pEqn;
  return nullptr;
}
phb_solver *phb_fs_solver(phb_fracstep *fs, const char *name) {
  if (!fs || !name) return nullptr;
  if (!strcmp(name, "uEqn")) return fs->uSolver;
  if (!strcmp(name, "pEqn")) return fs->pSolver;
  return nullptr;
}

// FractionalStep::initialize (US/FractionalStep.cpp:25-28)
int phb_fs_initialize(phb_fracstep *fs) {
  PHB_REQUIRE(fs, "phb_fs_initialize: NULL argument");
  PHB_CHECK(phb::field_send_messages(fs->u));
  PHB_CHECK(phb::field_interpolate_faces(fs->u));
  PHB_CHECK(phb::field_set_boundary_faces(fs->p));
  bool neumann = false;  // all-Neumann pressure: pEqn_ is singular, keep its right-hand side compatible
  PHB_CHECK(phb::field_all_neumann(fs->p, &neumann));
  PHB_CHECK(phb_solver_setup(fs->pSolver, "nullSpace", neumann ? "constant" : "none"));
  return PHB_OK;
}

// uEqn_ = (fv::ddt(u,dt) + fv::div(u,u,0.) == fv::laplacian(mu/rho,u,0.5) - src::src(gradP))
// (US/FractionalStep.cpp:82-83); expects u.savePreviousTimeStep to have run.
int phb_fs_assemble_u(phb_fracstep *fs, double dt) {
  PHB_REQUIRE(fs && dt > 0., "phb_fs_assemble_u: bad argument");
  fs->uTag = 0ull;
  if (fs->fusedAssembly) {
    const int rc = phb::assemble_momentum_predictor(fs->uEqn, fs->u, fs->gradP, fs->mu / fs->rho, dt);
    // the coefficients are V/dt and the halved diffusion links: a function of dt, nu and the boundary types alone
    if (rc == PHB_OK) fs->uTag = matrix_tag(1, dt, fs->mu / fs->rho, fs->u->bcVersion);
    if (rc <= 0) return rc;   // 1: SYMMETRY patches -> term by term
  }
  PHB_CHECK(phb_eqn_zero(fs->uEqn));
  PHB_CHECK(phb_assemble_ddt(fs->uEqn, fs->u, 1., nullptr, dt, +1.));
  PHB_CHECK(phb_assemble_div(fs->uEqn, fs->u, fs->u, 0., +1.));
  PHB_CHECK(phb_assemble_laplacian(fs->uEqn, fs->mu / fs->rho, nullptr, fs->u, 0.5, -1.));
  PHB_CHECK(phb_assemble_src(fs->uEqn, fs->gradP, +1.));  // == (... - src)  ->  rhs_ += src
  return PHB_OK;
}

// pEqn_ = (fv::laplacian(dt, p) == src::div(u))  (US/FractionalStep.cpp:97)
int phb_fs_assemble_p(phb_fracstep *fs, double dt) {
  PHB_REQUIRE(fs && dt > 0., "phb_fs_assemble_p: bad argument");
  fs->pTag = 0ull;
  if (fs->fusedAssembly) {
    PHB_CHECK(phb::assemble_pressure_poisson(fs->pEqn, fs->p, fs->u, dt));
    fs->pTag = matrix_tag(2, dt, 1., fs->p->bcVersion);   // coefficients dt g_f: dt and the boundary types
    return PHB_OK;
  }
  PHB_CHECK(phb_eqn_zero(fs->pEqn));
  PHB_CHECK(phb_assemble_laplacian(fs->pEqn, dt, nullptr, fs->p, -1., +1.));
  PHB_CHECK(phb_assemble_src_div(fs->pEqn, fs->u, -1.));
  return PHB_OK;
}

int phb_fs_step(phb_fracstep *fs, double dt, double stats[6]) {
  PHB_REQUIRE(fs && dt > 0., "phb_fs_step: bad argument");
  phb_ctx *c = fs->m->ctx;
  int itU = 0, itP = 0;
  double rrU = 0., rrP = 0.;
  // ---- solveUEqn (US/FractionalStep.cpp:79-94)
  PHB_CHECK(phb_field_save_previous(fs->u));
  PHB_CHECK(phb_fs_assemble_u(fs, dt));
  PHB_CHECK(phb::eqn_solve_tagged(fs->uEqn, fs->uSolver, fs->u, fs->warmStart, fs->uTag, &itU, &rrU));
  PHB_CHECK(phb::field_axpy_cells(fs->u, dt, fs->gradP));
  PHB_CHECK(phb::field_send_messages(fs->u));
  PHB_CHECK(phb::field_interpolate_faces(fs->u));
  // ---- solvePEqn (:96-107)
  PHB_CHECK(phb_fs_assemble_p(fs, dt));
  if (fs->warmStart && fs->guessOrder == 1) {
    // the reference passes no guess at all; p^n is the natural one, 2 p^n - p^(n-1) a better one
    const int n = fs->m->nLocal;
    PHB_CHECK(fs->pPrev.alloc((size_t)n));
    const int g = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, (long long)c->numSMs * 8));
    PHB_LAUNCH(c, k_extrapolate, g, 256, 0, n, fs->p->cells.p, fs->pPrev.p, fs->nStepsDone >= 2 ? 1 : 0);
  }
  PHB_CHECK(phb::eqn_solve_tagged(fs->pEqn, fs->pSolver, fs->p, fs->warmStart, fs->pTag, &itP, &rrP));
  fs->nStepsDone++;
  PHB_CHECK(phb::field_send_messages(fs->p));
  PHB_CHECK(phb::field_set_boundary_faces(fs->p));
  PHB_CHECK(phb::field_gradient(fs->p, fs->gradP));
  // ---- correctVelocity (:109-117)
  PHB_CHECK(phb::field_axpy_cells(fs->u, -dt, fs->gradP));
  PHB_CHECK(phb::field_send_messages(fs->u));
  PHB_CHECK(phb::field_axpy_faces(fs->u, -dt, fs->gradP));
  // ---- diagnostics (:41-43): max divergence error, max CFL
  PHB_CHECK(phb::field_flux_diagnostics(fs->u, dt, fs->scratch, fs->partials, fs->ticket, fs->out.p));
  PHB_CUDA(cudaMemcpyAsync(c->pinned + 64, fs->out.p, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  if (stats) {
    stats[0] = itU; stats[1] = itP; stats[2] = rrU; stats[3] = rrP;
    stats[4] = c->pinned[64]; stats[5] = c->pinned[65];
  }
  return phb::launch_status(c);
}

// State of a time step from CELL values alone -- what the reference itself persists: Solver::readLatestCgnsFlowSolution
// (US/Solver.cpp:544-581) reads the cell fields of a restart file and re-derives every face value.  Given the owned
// cells of u and p on the device and the time step dtPrev of the step that produced them, rebuild p's ghosts and
// boundary faces, gradP, and the face velocities exactly as phb_fs_step left them:
//   u* = u + dtPrev grad p (cells) -> interpolateFaces -> faces - dtPrev (grad p)_f    (FIXED boundary faces keep
// their boundary values), the cells themselves untouched.  dtPrev = 0 (state at rest / unknown): plain interpolation,
// the reference's own restart.
int phb_fs_rebuild_faces(phb_fracstep *fs, double dtPrev) {
  PHB_REQUIRE(fs && dtPrev >= 0., "phb_fs_rebuild_faces: bad argument");
  phb_ctx *c = fs->m->ctx;
  PHB_CHECK(phb::field_send_messages(fs->p));
  PHB_CHECK(phb::field_set_boundary_faces(fs->p));
  PHB_CHECK(phb::field_gradient(fs->p, fs->gradP));
  const size_t len = fs->u->cells.n;
  PHB_CHECK(fs->uSave.alloc(len));
  PHB_CUDA(cudaMemcpyAsync(fs->uSave.p, fs->u->cells.p, len * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  if (dtPrev > 0.) PHB_CHECK(phb::field_axpy_cells(fs->u, dtPrev, fs->gradP));
  PHB_CHECK(phb::field_send_messages(fs->u));
  PHB_CHECK(phb::field_interpolate_faces(fs->u));
  PHB_CUDA(cudaMemcpyAsync(fs->u->cells.p, fs->uSave.p, len * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  PHB_CHECK(phb::field_send_messages(fs->u));
  if (dtPrev > 0.) {
    PHB_CHECK(phb::field_face_types(fs->u));
    const long long nF = fs->m->nFaces;
    const int g = (int)std::max<long long>(1, std::min<long long>((nF + 255) / 256, (long long)c->numSMs * 8));
    PHB_LAUNCH(c, k_axpy_faces_nonfixed, g, 256, 0, nF, 2, (const int *)fs->u->dFaceType.p, -dtPrev,
               (const double *)fs->gradP->faces.p, fs->u->faces.p);
  }
  return phb::launch_status(c);
}

// driver options: "warmStart" 0/1 (guess = previous field values), "guessOrder" 0/1 (pEqn_ guess extrapolation)
int phb_fs_setup(phb_fracstep *fs, const char *key, double value) {
  PHB_REQUIRE(fs && key, "phb_fs_setup: NULL argument");
  if (!strcmp(key, "warmStart")) fs->warmStart = value != 0.;
  else if (!strcmp(key, "guessOrder")) fs->guessOrder = (int)value;
  else if (!strcmp(key, "fusedAssembly")) fs->fusedAssembly = value != 0.;
  else PHB_REQUIRE(false, "phb_fs_setup: unknown key \"%s\"", key);
  return PHB_OK;
}

// computeMaxTimeStep (US/FractionalStep.cpp:68-77)
int phb_fs_max_time_step(phb_fracstep *fs, double maxCo, double prevDt, double maxDt, double *out) {
  PHB_REQUIRE(fs && out && prevDt > 0., "phb_fs_max_time_step: bad argument");
  phb_ctx *c = fs->m->ctx;
  PHB_CHECK(phb::field_flux_max(fs->u, 1, prevDt, fs->scratch, fs->partials, fs->ticket, fs->out.p + 1));
  PHB_CUDA(cudaMemcpyAsync(c->pinned + 64, fs->out.p + 1, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  const double co = c->pinned[64];
  const double l1 = 0.1, l2 = 1.2;
  *out = std::fmin(std::fmin(maxCo / co * prevDt, (1 + l1 * maxCo / co) * prevDt), std::fmin(l2 * prevDt, maxDt));
  return PHB_OK;
}

}  // extern "C"
