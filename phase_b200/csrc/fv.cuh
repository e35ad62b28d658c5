// fv.cuh -- internal field/equation helpers shared by fv.cu and fracstep.cu.
#pragma once
#include "structs.cuh"

namespace phb {
int field_face_types(phb_field *f);
int field_interpolate_faces(phb_field *f);
int field_set_boundary_faces(phb_field *f);
int field_gradient(const phb_field *phi, phb_field *grad);
int field_axpy_cells(phb_field *y, double a, const phb_field *x);  // owned cells
int field_axpy_faces(phb_field *y, double a, const phb_field *x);  // all faces
int field_send_messages(phb_field *f);
int field_all_neumann(phb_field *f, bool *out);
// phb_eqn_solve with a coefficient tag (0 = none): equal tags = equal matrix, vouched for by the caller
int eqn_solve_tagged(phb_eqn *e, phb_solver *s, phb_field *phi, int warmStart, unsigned long long tag, int *iters,
                     double *relres);
// FractionalStep's uEqn_ (ddt + div == laplacian(gamma, theta 0.5) - src(gradP)) in one pass; 1 = not applicable
// FractionalStep's pEqn_ (laplacian(gamma, p) == src::div(u)) in one pass
int assemble_pressure_poisson(phb_eqn *e, phb_field *p, const phb_field *u, double gamma);
int assemble_momentum_predictor(phb_eqn *e, phb_field *u, const phb_field *gradP, double gamma, double dt);
// device max over owned cells of |sum_f u_f.S_f| (mode 0) or the Courant number (mode 1)
int field_flux_max(const phb_field *u, int mode, double dt, DevBuf<double> &scratch, DevBuf<double> &partials,
                   DevBuf<unsigned> &ticket, double *devOut);
// both at once: devOut[0] = max divergence error, devOut[1] = max Courant number
int field_flux_diagnostics(const phb_field *u, double dt, DevBuf<double> &scratch, DevBuf<double> &partials,
                           DevBuf<unsigned> &ticket, double *devOut);
}  // namespace phb
