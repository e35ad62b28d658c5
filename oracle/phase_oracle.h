/*
 * phase_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE ONLY).
 *
 * A flat-array C restatement of the reference's per-time-step linear-system
 * path (obrienadam/Phase, snapshot under /root/reference).  It exists to CHECK
 * the CUDA product in phase_b200/csrc and to serve as the timed CPU baseline in
 * bench.py.  Nothing in the product path may include, link or call this file.
 *
 * Parity status: PINNED on the reference's own code.
 *   (1) oracle/_ref/libphase_ref_fv.so: the reference's grid, field, equation,
 *       fv::/src::/cicsam:: operator and FractionalStep translation units compiled
 *       where they lie (oracle/build.py: build_ref_fv) over stand-in headers for
 *       Boost / MPI (one rank) / METIS / CGNS (oracle/ref_stub) and driven through
 *       oracle/ref_fv_driver.cpp.  tests/test_oracle_ref_fv.py: connectivity, link
 *       tables, IndexMap and the CSR patterns at the solver hand-off bit-exact;
 *       coefficients, right-hand sides and fields after K FractionalStep::solve
 *       calls to round-off.  tests/golden/ref_*.npz are written by that library
 *       (tests/golden/make_ref_golden.py) and checked against this file and the
 *       CUDA path wherever the tests run.
 *   (2) oracle/_ref/libphase_ref_crs.so: src/Math/{Vector,CrsEquation,
 *       SparseMatrixSolver}.cpp alone -- exact CSR insertion / compaction /
 *       operator algebra / set()+setRhs(-rhs) hand-off (tests/test_oracle_ref_crs.py).
 *   (3) hand-simulated known answers, tests/test_oracle_kat.py.
 *       The same library runs the reference's partition(), initCommBuffers() and
 *       IndexMap on several MPI ranks (threads, oracle/ref_mpi_threads.cpp) for a given
 *       partition vector: local meshes, ownership, buffer / send groups and global
 *       row numbers bit-exact (tests/test_oracle_ref_partition.py).
 * Not pinned: the polygon area / centroid arithmetic (Boost.Geometry is absent;
 * the stand-in implements the same shoelace / Bashein-Detmer formulas), the METIS
 * call that produces the partition vector (version / options unpinned: the vector
 * is an input) and the solve arithmetic itself, which lives in un-vendored
 * Eigen3 / Trilinos (versions unpinned): any exact direct solve (scipy splu)
 * stands in for it.
 */
#ifndef PHASE_ORACLE_H
#define PHASE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* boundary types: UF/FiniteVolumeField.h enum BoundaryType */
enum { OR_FIXED = 0, OR_NORMAL_GRADIENT = 1, OR_SYMMETRY = 2 };

typedef struct OrMesh OrMesh;
typedef struct OrCrs OrCrs;
typedef struct OrFracStep OrFracStep;

/* linear-solve callback used by the time-step drivers: solve A x = b, CSR with
 * -1 padded columns (exactly what SparseMatrixSolver::set receives). */
typedef int (*or_solve_cb)(int n, const int *rowPtr, const int *colInd,
                           const double *vals, const double *b, double *x,
                           void *user);

/* ---- mesh (I1, I2, G1-G4) ---- */
OrMesh *or_mesh_create(int nNodes, const double *xy, int nCells,
                       const int *cptr, const int *cind);
OrMesh *or_mesh_rectilinear(int nx, int ny, double width, double height);
/* each quad of an nx x ny grid split along alternating diagonals */
OrMesh *or_mesh_triangulated(int nx, int ny, double width, double height);
void or_mesh_destroy(OrMesh *m);
/* patches by node pairs (FiniteVolumeGrid2D::createPatchByNodes) */
int or_mesh_add_patch_by_nodes(OrMesh *m, const char *name, int nPairs,
                               const int *nodePairs);
int or_mesh_patch_id(const OrMesh *m, const char *name);
/* generic array accessor: returns length, sets *ptr and *isDouble */
long or_mesh_array(const OrMesh *m, const char *name, const void **ptr,
                   int *isDouble);

/* ---- partition + halo maps (I3, I5) ---- */
/* local mesh of `rank` given a global mesh and a cell partition vector */
OrMesh *or_mesh_partition_local(const OrMesh *g, const int *cellPartition,
                                int rank, int nProcs);
/* after all local meshes exist: build send lists (initCommBuffers) */
int or_mesh_init_comm(OrMesh **locals, int nProcs);

/* ---- CrsEquation restatement (I4, A8) ---- */
OrCrs *or_crs_create(int nRows, int nnzPerRow);
OrCrs *or_crs_clone(const OrCrs *e);
void or_crs_destroy(OrCrs *e);
void or_crs_add_coeff(OrCrs *e, int row, int col, double v);
void or_crs_set_coeff(OrCrs *e, int row, int col, double v);
void or_crs_add_rhs(OrCrs *e, int row, double v);
void or_crs_scale_row(OrCrs *e, int row, double v);
void or_crs_add_eq(OrCrs *lhs, const OrCrs *rhs);  /* operator+= */
void or_crs_sub_eq(OrCrs *lhs, const OrCrs *rhs);  /* operator-=, operator== */
void or_crs_sub_vec(OrCrs *lhs, const double *v);  /* operator-=(Vector) */
void or_crs_add_vec(OrCrs *lhs, const double *v);
void or_crs_scale(OrCrs *lhs, double s);
int or_crs_rank(const OrCrs *e);
int or_crs_nnz(const OrCrs *e); /* stored slots incl. padding */
void or_crs_export(const OrCrs *e, int *rowPtr, int *colInd, double *vals,
                   double *rhs);

/* ---- fractional-step driver (A2-A7, A10, S1) ---- */
OrFracStep *or_fs_create(OrMesh *m, double rho, double mu);
void or_fs_destroy(OrFracStep *s);
/* boundary condition of field "u" or "p" on a patch */
int or_fs_set_bc(OrFracStep *s, const char *field, const char *patch, int type,
                 double vx, double vy);
void or_fs_initialize(OrFracStep *s);
void or_fs_set_solver(OrFracStep *s, or_solve_cb cb, void *user);
/* parameters of the built-in BiCGStab used when no callback is set */
void or_fs_set_solver_params(OrFracStep *s, double tol, int maxIters,
                             int precond);
/* one FractionalStep::solve(dt); returns 0 */
int or_fs_step(OrFracStep *s, double dt);
/* only assemble (no solve): for assembly parity + CPU assembly timing */
void or_fs_assemble_u(OrFracStep *s, double dt);
void or_fs_assemble_p(OrFracStep *s, double dt);
const OrCrs *or_fs_ueqn(const OrFracStep *s);
const OrCrs *or_fs_peqn(const OrFracStep *s);
long or_fs_array(OrFracStep *s, const char *name, double **ptr);
double or_fs_max_divergence(const OrFracStep *s);
double or_fs_max_courant(OrFracStep *s, double dt);
int or_fs_last_iters(const OrFracStep *s, int which);

/* ---- stand-alone operators for operator-level parity ---- */
/* variable-coefficient Poisson: laplacian(Field gamma, p) == div(u)
 * (FractionalStepMultiphase::solvePEqn), gammaCell/gammaFace given. */
OrCrs *or_op_laplacian_field(OrFracStep *s, const double *gammaFace);

/* uEqn_ of FractionalStepMultiphase (rho*ddt + rho*dive == laplacian(mu) + src(f)) */
OrCrs *or_op_ueqn_multiphase(OrFracStep *s, double dt, const double *rhoCell,
                             const double *muFace, const double *mu0Face,
                             const double *fx, const double *fy);
/* (fv::ddt(rho, phi, dt) + fv::div(u, phi, theta) == 0) on the scalar field "p" */
OrCrs *or_op_scalar_transport(OrFracStep *s, double dt, double theta, const double *rho,
                              const double *rho0, const double *phi0, const double *phi0f);

/* CICSAM (A9): UD/Cicsam.cpp.  gamma = the scalar field "p", u = faces of "u". */
void or_cicsam_weights(OrFracStep *s, double dt, const double *gradGx, const double *gradGy, double *beta);
OrCrs *or_op_cicsam_div(OrFracStep *s, double theta, const double *beta, const double *gamma0, const double *gamma0f);
void or_cicsam_momentum_flux(OrFracStep *s, double rho1, double rho2, const double *beta, double *outx, double *outy);

/* ---- built-in CPU solver (OpenMP BiCGStab, Jacobi or ILU(0)) ---- */
/* precond: 0 none, 1 Jacobi, 2 ILU(0).  returns iterations, *relres out.
 * x holds the initial guess on entry. */
int or_bicgstab(int n, const int *rowPtr, const int *colInd,
                const double *vals, const double *b, double *x, double tol,
                int maxIters, int precond, double *relres);
/* multicolour ordering (the CUDA path's rule) and BiCGStab + ILU(0) on a system
 * grouped in independent sets: triangular sweeps OpenMP-parallel inside a set */
int or_multicolor_order(int n, const int *rowPtr, const int *colInd, int *new2old,
                        int *blockPtr, int maxColours);
int or_bicgstab_blocks(int n, const int *rowPtr, const int *colInd, const double *vals,
                       const double *b, double *x, double tol, int maxIters,
                       int nBlocks, const int *blockPtr, double *relres);
int or_num_threads(void);
int or_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
