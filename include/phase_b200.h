/*
 * phase_b200.h -- C ABI of libphase_b200.so: the B200-native implementation of
 * Phase's per-time-step linear-system path (fv:: operator assembly into a CSR
 * FiniteVolumeEquation, then the BiCGStab solve), hand-written CUDA for sm_100a.
 *
 * Conventions
 *   - every function returns 0 on success, a negative phb_status on failure and
 *     never throws; phb_last_error() gives the text (thread-local).
 *   - every pointer argument is CALLER-OWNED HOST memory unless the name says
 *     "device"; calls are synchronous from one host thread per context.
 *   - Scalar = double, Index = int32 (reference: src/Types/Types.h:7-10).
 *   - an equation is stored as  A x + rhs = 0  and the solver receives b = -rhs
 *     (reference: UE/FiniteVolumeEquation.tpp:71-73, M/CrsEquation.cpp:169-175).
 *
 * Reference paths below are relative to /root/reference/src; UG = 2D/Unstructured/
 * FiniteVolumeGrid2D, UF = 2D/Unstructured/FiniteVolume/Field, UD = .../Discretization,
 * UE = .../Equation, US = 2D/Unstructured/Solvers, M = Math.
 */
#ifndef PHASE_B200_H
#define PHASE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  PHB_OK = 0,
  PHB_ERR_ARG = -1,        /* bad argument / unknown key / size mismatch */
  PHB_ERR_CUDA = -2,       /* CUDA runtime error */
  PHB_ERR_COMM = -3,       /* NCCL error */
  PHB_ERR_BREAKDOWN = -4,  /* Krylov breakdown (rho or omega = 0, non-finite) */
  PHB_ERR_NOT_CONVERGED = -5,
  PHB_ERR_UNSUPPORTED = -6,
  PHB_ERR_STATE = -7       /* call order (e.g. solve before set) */
} phb_status;

/* boundary types: UF/FiniteVolumeField.h:12 enum BoundaryType */
enum { PHB_FIXED = 0, PHB_NORMAL_GRADIENT = 1, PHB_SYMMETRY = 2 };
/* preconditioners */
enum { PHB_PC_NONE = 0, PHB_PC_JACOBI = 1, PHB_PC_ILU0 = 2, PHB_PC_AMG = 3 };

typedef struct phb_ctx phb_ctx;
typedef struct phb_mesh phb_mesh;
typedef struct phb_solver phb_solver;
typedef struct phb_field phb_field;
typedef struct phb_eqn phb_eqn;
typedef struct phb_fracstep phb_fracstep;

const char *phb_last_error(void);
int phb_version(void);

/* ------------------------------------------------------------------ context
 * One context = one GPU + its streams (+ an NCCL communicator for nProcs>1).
 * Replaces: S/Communicator.{h,cpp} (MPI_Init, rank/nProcs, point-to-point and
 * all-reduce used around the solve, S/Communicator.cpp:11-17,81-141). */
/* device >= 0: a CUDA device (fails when none is present: there is no CPU
 * fallback).  PHB_DEVICE_HOST_ONLY: a context that can only build meshes,
 * partitions and halo maps on the host (integer artefacts I1-I5); creating a
 * field, equation or solver on it fails with PHB_ERR_STATE. */
#define PHB_DEVICE_HOST_ONLY (-1)
int phb_ctx_create(int device, phb_ctx **out);
int phb_ctx_destroy(phb_ctx *ctx);
/* 128-byte NCCL unique id made on rank 0, shipped by the host launcher */
int phb_comm_unique_id(void *out128);
int phb_ctx_init_comm(phb_ctx *ctx, int rank, int nProcs, const void *id128);
/* Peer-memory communication (NVLink, CUDA IPC) for the exchanges INSIDE the
 * Krylov loop: every rank allocates one arena (sized for `maxSolvers` solvers with
 * vectors of up to 2 x maxCols doubles; maxCols = max over ranks of owned+ghost
 * cells), the launcher all-gathers the 64-byte handles, and every rank opens the
 * others'.  After that halos and dot-product reductions of a distributed solve are
 * single small kernels storing into the peers' arenas; NCCL remains the path for
 * everything outside the loop.  Optional: without it every exchange uses NCCL. */
int phb_ctx_peer_arena_create(phb_ctx *ctx, long long maxCols, int maxSolvers,
                              void *handleOut64);
int phb_ctx_peer_arena_open(phb_ctx *ctx, const void *handles /* nProcs x 64 B */);
int phb_ctx_rank(const phb_ctx *ctx);
int phb_ctx_nprocs(const phb_ctx *ctx);
int phb_ctx_sync(phb_ctx *ctx);
/* number of kernels this library launched on the context so far */
long long phb_ctx_kernel_launches(const phb_ctx *ctx);
/* raw stream handle (cudaStream_t) the kernels run on, for CUDA-event timing */
void *phb_ctx_stream(phb_ctx *ctx);

/* --------------------------------------------------------------------- mesh
 * Host-side build of connectivity + geometry, flattened to SoA and uploaded.
 * Replaces: FiniteVolumeGrid2D::init(nodes,cptr,cind) + createCell + init()
 * (UG/FiniteVolumeGrid2D.cpp:20-35,84-114,396-450), Face/Cell/Link geometry
 * (UG/Face/Face.cpp:9-18,48-64; UG/Cell/Cell.cpp:8-29; UG/Link/ *.cpp),
 * StructuredRectilinearGrid::init/initPatches (UG/StructuredRectilinearGrid.cpp:40-95,175-194),
 * createPatchByNodes (UG/FiniteVolumeGrid2D.cpp:182-198). */
int phb_mesh_create(phb_ctx *ctx, int nNodes, const double *xy, int nCells,
                    const int *cptr, const int *cind, phb_mesh **out);
int phb_mesh_create_rectilinear(phb_ctx *ctx, int nx, int ny, double width,
                                double height, phb_mesh **out);
/* synthetic unstructured variant: every quad split along alternating diagonals */
int phb_mesh_create_triangulated(phb_ctx *ctx, int nx, int ny, double width,
                                 double height, phb_mesh **out);
int phb_mesh_add_patch_by_nodes(phb_mesh *m, const char *name, int nPairs,
                                const int *nodePairs);
int phb_mesh_patch_id(const phb_mesh *m, const char *name);
/* build links, canonical CSR pattern, face->slot map; upload.  Must be called
 * once after the patches are defined and before any field/equation is made. */
int phb_mesh_finalize(phb_mesh *m);
int phb_mesh_destroy(phb_mesh *m);
/* sizes: [nNodes, nCells, nFaces, nPatches, rank, nProcs, nLocal, rowOffset,
 *         nInteriorFaces, nBoundaryFaces, nnzScalar] */
int phb_mesh_sizes(const phb_mesh *m, long long out[11]);
/* host copies of mesh artefacts by name, for bit-exact parity checks.
 * int arrays:  cptr cind faceN1 faceN2 faceL faceR facePatch ilPtr ilFace ilCell
 *              blPtr blFace dlPtr dlCell rowPtr colInd slotL slotR slotDiag
 *              owner globalId localRow globalRow bufPtr bufCell sendPtr sendCell
 * f64 arrays:  vol cellCx cellCy faceCx faceCy faceSx faceSy faceG faceW
 * returns the length (or <0); copies min(length, cap) entries when out != NULL */
long long phb_mesh_get_i32(const phb_mesh *m, const char *name, int *out,
                           long long cap);
long long phb_mesh_get_f64(const phb_mesh *m, const char *name, double *out,
                           long long cap);

/* partition (I5): UG/FiniteVolumeGrid2D.cpp:276-392 with the partition VECTOR as
 * an input (the reference gets it from METIS_PartMeshDual, :287-297).
 * phb_partition_rcb: deterministic recursive coordinate bisection of cell
 * centroids into nParts (any nParts >= 1). */
int phb_partition_rcb(const phb_mesh *global, int nParts, int *cellPartition);
/* METIS, as the reference calls it: method 0 = METIS_PartMeshDual with ncommon 2 (FiniteVolumeGrid2D::partition,
 * UG/FiniteVolumeGrid2D.cpp:287-297), method 1 = METIS_PartGraphRecursive on the face-neighbour graph (the
 * PhasePartitionGrid utility, U/utilities/PhasePartitionGrid.cpp:42-50).  PHB_ERR_UNSUPPORTED when the library was
 * built without the toolkit's libmetis_static.a.  objective (may be NULL) = edge cut reported by METIS. */
int phb_partition_metis(const phb_mesh *global, int nParts, int method, int *cellPartition, long long *objective);
/* What PhasePartitionGrid writes per partition (U/utilities/PhasePartitionGrid.cpp:56-153): cells = owned cells in
 * ascending id, then the halo cells in discovery order (cellLinks of the partition-boundary cells, then the cells
 * within minBufferWidth of them); GlobalID / ProcNo per cell; nodes renumbered in first-use order; element lists and
 * patch node pairs with 1-based local node ids.  tools/partition_grid.py writes them as ADF-CGNS files. */
typedef struct phb_partfile phb_partfile;
int phb_partition_file_build(const phb_mesh *global, const int *cellPartition, int proc, double minBufferWidth,
                             phb_partfile **out);
int phb_partition_file_sizes(const phb_partfile *f, long long out[4]); /* nCells nNodes len(eind) nPatches */
int phb_partition_file_get(const phb_partfile *f, int *globalId, int *procNo, double *nodesXY, int *eptr, int *eind);
long long phb_partition_file_patch(const phb_partfile *f, int p, char *name, int cap, int *nodePairs);
int phb_partition_file_destroy(phb_partfile *f);
/* local mesh of ctx's rank: owned cells + every face- or node-neighbour of an
 * owned cell (buffer layer), reference numbering (ascending global id), halo
 * lists as in initCommBuffers (:460-511).  `global` must be finalized. */
int phb_mesh_create_local(phb_ctx *ctx, const phb_mesh *global,
                          const int *cellPartition, phb_mesh **out);

/* ingest.  ADF-format CGNS file with the semantics of
 * CgnsUnstructuredGrid::load (UG/CgnsUnstructuredGrid.cpp:13-105): element sections
 * merged by element id, BC point lists = BAR_2 element ids -> patches by node pair,
 * cells = elements with > 2 nodes in element-id order.  The mesh is NOT finalized. */
int phb_mesh_read_cgns(phb_ctx *ctx, const char *filename, phb_mesh **out);
/* uniform refinement (tri -> 4 tri, quad -> 4 quad), `levels` rounds; patches follow */
int phb_mesh_refine(phb_ctx *ctx, const phb_mesh *in, int levels, phb_mesh **out);
int phb_mesh_patch_name(const phb_mesh *m, int id, char *out, int cap);

/* peer-memory halo layout: for every peer q, the offset at which this rank's
 * values land in q's vectors (q's "recvOff"[this rank]) and q's owned+ghost count */
int phb_mesh_set_peer_layout(phb_mesh *m, const int *peerRecvOff, const int *peerLd);

/* local mesh of ctx's rank for a y-strip partition of an nx x ny rectilinear
 * grid (rank r owns rows [r ny/P, (r+1) ny/P)) built without the global mesh;
 * same result as create_rectilinear + that partition + create_local. */
int phb_mesh_create_rect_strip(phb_ctx *ctx, int nx, int ny, double width,
                               double height, phb_mesh **out);
/* same for a px x py block partition (rank q = bj*px + bi); px*py = nProcs */
int phb_mesh_create_rect_block(phb_ctx *ctx, int nx, int ny, double width,
                               double height, int px, int py, phb_mesh **out);

/* ------------------------------------------------------------ linear solver
 * Seam 1.  Beneath class SparseMatrixSolver (M/SparseMatrixSolver.h:11-62):
 *   setup(ptree LinearAlgebra.<eqn>)  -> phb_solver_setup(key,value) per key
 *        keys (M/TrilinosBelosSparseMatrixSolver.cpp:44-86): solver, maxIters,
 *        tolerance, preconditioner {none,jacobi,ilu0,schwarz}, iluFill
 *   setRank                            -> phb_solver_set_rank
 *   set(rowPtr,colInds,vals)           -> phb_solver_set_csr   (cols global, -1 = padding)
 *   set(vector<SparseEntry>) / tuples  -> phb_solver_set_coo   (duplicates summed, M/SparseMatrixSolver.cpp:5-31)
 *   setRhs / setGuess                  -> phb_solver_set_rhs / phb_solver_set_guess
 *   solve / nIters / error             -> phb_solver_solve
 *   x(i)                               -> phb_solver_get_x (whole vector, host)
 * BiCGStab, right-preconditioned (Jacobi folded into the matrix, or level-
 * scheduled ILU(0)); convergence ||r||_2 <= tolerance * ||b||_2. */
int phb_solver_create(phb_ctx *ctx, phb_solver **out);
int phb_solver_destroy(phb_solver *s);
int phb_solver_setup(phb_solver *s, const char *key, const char *value);
int phb_solver_set_rank(phb_solver *s, int nRows, int nCols);
int phb_solver_set_csr(phb_solver *s, int nRows, const int *rowPtr,
                       const int *colInd, const double *vals);
int phb_solver_set_coo(phb_solver *s, int nRows, long long nEntries,
                       const int *rows, const int *cols, const double *vals);
/* distributed systems: owned rows are [rowOffset, rowOffset+nRows) of the
 * global numbering; ghost columns are resolved through the mesh halo lists */
int phb_solver_set_halo(phb_solver *s, const phb_mesh *localMesh, int nComp);
int phb_solver_set_rhs(phb_solver *s, const double *b, int n);
int phb_solver_set_guess(phb_solver *s, const double *x0, int n);
int phb_solver_solve(phb_solver *s, int *iters, double *relres);
int phb_solver_get_x(const phb_solver *s, double *x, int n);
/* y = A x on the device copy of the last matrix set (host in/out): SpMV parity */
int phb_solver_spmv(phb_solver *s, const double *x, double *y, int n);
/* repeat the device SpMV `reps` times on resident data; returns mean ms/launch
 * measured with CUDA events on the context stream (bench.py roofline leg) */
int phb_solver_time_spmv(phb_solver *s, int reps, double *msPerLaunch);
/* z = M^-1 r with the multigrid hierarchy of the last solve (host vectors over the owned rows, [x-block | y-block]):
 * lets tests compare one application of the device V-cycle with a transcription of the same hierarchy */
int phb_solver_apply_preconditioner(phb_solver *s, const double *r, double *z, int n);
/* algorithmic byte counts of the current matrix: [spmv, bicgstabIteration] */
int phb_solver_bytes(const phb_solver *s, double out[2]);

/* `preconditioner amg` -- smoothed-aggregation multigrid V-cycle (the reference's `lib muelu`,
 * Math/TrilinosMueluSparseMatrixSolver.cpp:27-32).  Extra setup keys: amgTheta (strength threshold, 0),
 * amgCoarsest (rows of the densely inverted coarsest level, at most 1000 = default), amgSweeps (Jacobi sweeps before and
 * after the coarse correction, 1), amgSmootherWeight (1.8, divided by the Gershgorin bound of
 * rho(D^-1 A); level 0), amgCoarseSmootherWeight (1.6, divided by a power-iteration estimate of lambda_max(D^-1 A);
 * Galerkin levels; 0 = Gershgorin rule everywhere), amgAggTheta (0.1: couplings below this fraction of the row's largest
 * do not make a neighbour a member of the row's aggregate), amgRebuild (auto | always).  The hierarchy is built on the host once per matrix and
 * reused while the matrix stays a scalar multiple of it or the iteration count does not degrade.
 * info = [levels, operator complexity, host setup ms, setups so far, coarsest rows, kernel launches
 *         per cycle, iterations of the first solve after the last setup, hierarchy stale (0/1)] */
int phb_solver_amg_info(const phb_solver *s, double info[8]);
/* Numeric re-setup on the device (key `amgRefresh off | auto | always`, default auto; single-rank hierarchies): when the
 * coefficients change on the same pattern -- pEqn_ = laplacian(dt / rho, p) of FractionalStepMultiphase
 * (US/FractionalStepMultiphase.cpp:129-148), every uEqn_ -- the aggregates, strength flags and the patterns of
 * P, R, A P, R A P of the last host setup are kept and the VALUES of every level (smoothed prolongators, Galerkin
 * products, smoother weights, dense coarsest inverse) are recomputed by kernels from the resident matrix.  `auto`
 * does so once the iteration count has drifted 20 % above the count after the setup; a refresh that does not bring it
 * back triggers the host setup again.  phb_solver_amg_refresh forces one now.
 * info = [refreshes since the last host setup, device ms of the last one, iterations of the first solve after it,
 *         symbolic data resident (0/1), its bytes, 0, 0, 0] */
int phb_solver_amg_refresh(phb_solver *s);
int phb_solver_amg_refresh_info(const phb_solver *s, double info[8]);
/* Values of one matrix of the hierarchy as the cycle streams them (sliced-ELL slot order, as double): which = 0
 * operator (levels >= 1), 1 prolongator, 2 restriction, 3 smoother weights, 4 dense coarsest inverse.  Returns the
 * count (out == NULL: only the count) or a negative error.  Test hook of the numeric re-setup. */
long long phb_solver_amg_values(const phb_solver *s, int level, int which, double *out, long long cap);
/* live timing (CUDA events, resident data) of the cycle's level-0 kernels and of one whole cycle:
 * out = ms per launch of [residual, restriction, prolongation, Jacobi sweep], ms per cycle, algorithmic bytes of
 * the Jacobi launch, of the cycle, launches per cycle (bench.py roofline leg; single rank) */
int phb_solver_time_amg(phb_solver *s, int reps, double out[8]);
/* The same setup on a host CSR matrix, level matrices readable (works on a host-only context; used by
 * the CPU tests to check the Galerkin products and the cycle against scipy).
 * which: 0 = A_l, 1 = P_l (n_l x n_{l+1}), 2 = R_l = P_l^T */
typedef struct phb_amg_host phb_amg_host;
int phb_amg_host_build(int n, const int *rowPtr, const int *colInd, const double *vals, double theta,
                       int coarsest, phb_amg_host **out);
/* the same with the two setup rules of round 2 exposed: aggTheta (`amgAggTheta`) and coarseWeight
 * (`amgCoarseSmootherWeight`); a negative value selects the default (0.1, 1.6), 0 switches the rule off */
int phb_amg_host_build_ex(int n, const int *rowPtr, const int *colInd, const double *vals, double theta,
                          double aggTheta, double coarseWeight, int coarsest, phb_amg_host **out);
int phb_amg_host_levels(const phb_amg_host *h, int *nLevels, int *singular, int *denseCoarse);
int phb_amg_host_level_size(const phb_amg_host *h, int level, int which, int *nRows, int *nCols,
                            long long *nnz, double *rho);
int phb_amg_host_level_csr(const phb_amg_host *h, int level, int which, int *rowPtr, int *colInd,
                           double *vals);
/* smoother weights of a level = wScale / a_ii (Gershgorin rule on level 0, power-iteration rule on Galerkin levels) */
int phb_amg_host_level_weight(const phb_amg_host *h, int level, double *wScale);
int phb_amg_host_coarse_inverse(const phb_amg_host *h, double *inv);
int phb_amg_host_destroy(phb_amg_host *h);
/* The distributed setup used when nProcs > 1 (`amgScope global`): aggregates are rank-local, the prolongator
 * is smoothed across the rank boundaries, the Galerkin operators R A P keep the couplings between ranks, the level with <= tailRows global rows is gathered and the
 * rest of the hierarchy replicated.  Test hook: the ranks run as threads of this process, `part[i]` is
 * the owner of global row i.  Matrices come back with GLOBAL ids of their level (which: 0 = A_l, 1 = P_l,
 * 2 = R_l: the restriction is the part of P_l^T inside the rank, so it needs no communication). */
typedef struct phb_amg_dist phb_amg_dist;
int phb_amg_dist_build(int nRanks, int n, const int *rowPtr, const int *colInd, const double *vals,
                       const int *part, double theta, int coarsest, long long tailRows, phb_amg_dist **out);
int phb_amg_dist_info(const phb_amg_dist *h, int *nDistLevels, int *nTailLevels, int *singular);
const phb_amg_host *phb_amg_dist_tail(const phb_amg_dist *h, int rank);
int phb_amg_dist_matrix_size(const phb_amg_dist *h, int rank, int level, int which, int *nRows, long long *nnz);
int phb_amg_dist_matrix(const phb_amg_dist *h, int rank, int level, int which, int *rowPtr, int *colGid,
                        double *vals, int *rowGid);
int phb_amg_dist_halo(const phb_amg_dist *h, int rank, int level, int *sendPtr, int *sendIdx, int *recvPtr);
int phb_amg_dist_ghost_gids(const phb_amg_dist *h, int rank, int level, int *out);
/* which + 10 in phb_amg_dist_matrix*: the rank's LOCAL column numbering (what the device cycle works on) */
int phb_amg_dist_tail_offsets(const phb_amg_dist *h, int *offsets);
int phb_amg_dist_destroy(phb_amg_dist *h);

/* ---------------------------------------------------------- fields, equations
 * Seam 2.  Device mirrors of FiniteVolumeField<T> (cells + faces, BC table,
 * one history level) and of FiniteVolumeEquation<T> on the canonical pattern
 * row P = [P, nb in link order].
 * Replaces: UF/FiniteVolumeField.{h,tpp}, UE/FiniteVolumeEquation.{h,tpp},
 * UE/{Scalar,Vector}FiniteVolumeEquation.cpp. */
int phb_field_create(phb_mesh *m, int nComp, const char *name, phb_field **out);
int phb_field_destroy(phb_field *f);
int phb_field_set_bc(phb_field *f, const char *patch, int type, double vx, double vy);
/* part: "cells" | "faces" | "cells0" | "faces0" (old time level); host arrays
 * are component-blocked: [x-block | y-block] */
int phb_field_set(phb_field *f, const char *part, const double *v, long long n);
int phb_field_get(const phb_field *f, const char *part, double *v, long long n);
int phb_field_fill(phb_field *f, double vx, double vy);
/* savePreviousTimeStep(dt,1): UF/FiniteVolumeField.tpp:208-227 */
int phb_field_save_previous(phb_field *f);
/* interpolateFaces(DISTANCE) + setBoundaryFaces: UF/FiniteVolumeField.tpp:129-182,
 * UF/VectorFiniteVolumeField.cpp:140-161 */
int phb_field_interpolate_faces(phb_field *f);
int phb_field_set_boundary_faces(phb_field *f);
/* ScalarGradient::compute(FACE_TO_CELL): UF/ScalarGradient.cpp:34-74 */
int phb_field_gradient(const phb_field *phi, phb_field *grad);
/* halo exchange of the cell values (grid_->sendMessages(field)),
 * UG/FiniteVolumeGrid2D.tpp:3-49 */
int phb_field_send_messages(phb_field *f);

int phb_eqn_create(phb_mesh *m, int nComp, phb_eqn **out);
int phb_eqn_destroy(phb_eqn *e);
int phb_eqn_zero(phb_eqn *e);
/* every assemble call ACCUMULATES sign * operator into e (sign = +1 for the
 * left-hand side of `==`, -1 for the right-hand side):
 *   ddt        UD/TimeDerivative.h:7-48      (rhoField may be NULL -> rhoConst)
 *   div        UD/Divergence.h:8-53          (upwind, theta-weighted)
 *   dive       UD/ExplicitDivergence.h:7-51
 *   laplacian  UD/Laplacian.h:7-167, UD/Laplacian.cpp:5-119 (gammaField NULL -> gammaConst;
 *              theta < 0 selects the steady overloads without old-time terms)
 *   src        UD/Source.cpp:77-95           rhs += sign * field * V
 *   src_div    UD/Source.cpp:5-25            rhs += sign * sum_f u_f . S_f  */
int phb_assemble_ddt(phb_eqn *e, const phb_field *phi, double rhoConst,
                     const phb_field *rhoField, double dt, double sign);
int phb_assemble_div(phb_eqn *e, const phb_field *u, const phb_field *phi,
                     double theta, double sign);
int phb_assemble_dive(phb_eqn *e, const phb_field *u, const phb_field *phi,
                      double theta, double sign);
int phb_assemble_laplacian(phb_eqn *e, double gammaConst,
                           const phb_field *gammaField, const phb_field *phi,
                           double theta, double sign);
int phb_assemble_src(phb_eqn *e, const phb_field *f, double sign);
int phb_assemble_src_div(phb_eqn *e, const phb_field *u, double sign);
/* cell-group overloads (the immersed-boundary modules assemble on a CellGroup): fv::ddt(field, dt, cells)
 * (UD/TimeDerivative.h:50-62) and src::div(field, cells) (UD/Source.cpp:5-21); `cells` = reference cell ids */
int phb_assemble_ddt_cells(phb_eqn *e, const phb_field *phi, double dt, double sign, int nCells, const int *cells);
int phb_assemble_src_div_cells(phb_eqn *e, const phb_field *u, double sign, int nCells, const int *cells);
/* src::laplacian (UD/Source.cpp:27-75): rhs += sign * sum_links Gamma g_f (phi_nb - phi_P) (boundary links: phi_f).
 * gammaField NULL: scalar Gamma.  gammaField given: Gamma_f from the field's faces (the reference's field overload,
 * :50-75, indexes a one-component index map out of bounds and cannot be run; its evident intent is built). */
int phb_assemble_src_laplacian(phb_eqn *e, double gammaConst, const phb_field *gammaField, const phb_field *phi,
                               double sign);
/* CICSAM (UD/Cicsam.cpp): face weights beta_f (:19-66) into beta's FACE values,
 * cicsam::div(u, gamma, beta, theta) (:89-138) accumulated with sign, and the
 * density-weighted momentum flux rhoU_f (:69-87) into rhoU's face values. */
int phb_cicsam_weights(const phb_field *u, const phb_field *gamma,
                       const phb_field *gradGamma, double dt, phb_field *beta);
int phb_assemble_cicsam_div(phb_eqn *e, const phb_field *u, const phb_field *gamma,
                            const phb_field *beta, double theta, double sign);
int phb_cicsam_momentum_flux(double rho1, double rho2, const phb_field *u,
                             const phb_field *gamma, const phb_field *beta,
                             phb_field *rhoU);
/* rho * eqn row scaling: UE/VectorFiniteVolumeEquation.cpp:163-170 */
int phb_eqn_scale_rows(phb_eqn *e, const phb_field *rho);
/* relax(omega): body recovered from UE/ScalarFiniteVolumeEquation.cpp:45-55 */
int phb_eqn_relax(phb_eqn *e, const phb_field *phi, double omega);
/* host copy in the REFERENCE layout (I4), for parity:
 *   layout 0: compact [P, nb...] with exact zeros dropped (any sum of operators,
 *             M/CrsEquation.cpp:185-275); vector equations as 2N rows, y block offset N
 *   layout 1: ELL-5 padded [nb0, P, nb1, ...,-1] (fv::laplacian(Scalar,phi), UD/Laplacian.h:49-83)
 *   layout 2: ELL-5 padded [P, nb0, ...,-1]      (fv::laplacian(Field,phi),  UD/Laplacian.h:132-167)
 * pass NULLs to query sizes: returns nnz (stored slots incl. padding) */
long long phb_eqn_export_csr(const phb_eqn *e, int layout, int *rowPtr,
                             int *colInd, double *vals, double *rhs);
/* FiniteVolumeEquation<T>::solve: UE/FiniteVolumeEquation.tpp:64-86 -- device
 * resident (no host round trip); the solution is written into phi's cells;
 * phi's current cells are the initial guess when warmStart != 0 */
int phb_eqn_solve(phb_eqn *e, phb_solver *s, phb_field *phi, int warmStart,
                  int *iters, double *relres);

/* ------------------------------------------------- fractional-step time step
 * Device-resident FractionalStep::solve (US/FractionalStep.cpp:36-135): the
 * caller of the hot path for configs 1-3, built from the calls above. */
int phb_fs_create(phb_mesh *m, double rho, double mu, phb_fracstep **out);
int phb_fs_destroy(phb_fracstep *fs);
phb_field *phb_fs_field(phb_fracstep *fs, const char *name); /* u p gradP */
phb_eqn *phb_fs_eqn(phb_fracstep *fs, const char *name);      /* uEqn pEqn */
phb_solver *phb_fs_solver(phb_fracstep *fs, const char *name);
/* options: "warmStart" (1: solves start from the current field values; the reference
 * passes no guess), "guessOrder" (1: pEqn_ guess = 2 p^n - p^(n-1)), "fusedAssembly" (1, default: uEqn_ and
 * pEqn_ each assembled in one pass over the rows; 0: one kernel per fv:: / src:: operator, the path the operator
 * interface phb_assemble_* takes -- same coefficients to rounding) */
int phb_fs_setup(phb_fracstep *fs, const char *key, double value);
int phb_fs_initialize(phb_fracstep *fs);
/* A step's state from CELL values alone, as the reference's restart does (Solver::readLatestCgnsFlowSolution,
 * US/Solver.cpp:544-581: cell fields are read, faces re-derived): with the owned cells of u and p on the device and
 * dtPrev = the time step of the step that produced them, rebuilds ghosts, boundary faces, gradP and the face
 * velocities as phb_fs_step left them (dtPrev = 0: plain interpolateFaces, the reference's restart). */
int phb_fs_rebuild_faces(phb_fracstep *fs, double dtPrev);
int phb_fs_assemble_u(phb_fracstep *fs, double dt);
int phb_fs_assemble_p(phb_fracstep *fs, double dt);
/* stats: [itersU, itersP, relresU, relresP, maxDivergence, maxCourant] */
int phb_fs_step(phb_fracstep *fs, double dt, double stats[6]);
/* computeMaxTimeStep: US/FractionalStep.cpp:68-77 */
int phb_fs_max_time_step(phb_fracstep *fs, double maxCo, double prevDt,
                         double maxDt, double *out);

/* ------------------------------------------------- multiphase fractional step
 * Device-resident FractionalStepMultiphase::solve (US/FractionalStepMultiphase.cpp:52-218), config 4: CICSAM
 * advection of gamma, density / viscosity blend, gravity and CELESTE surface-tension sources
 * (U/FiniteVolume/Multiphase/{Celeste,CelesteStencil,SurfaceTensionForce,SurfaceTensionForceSmoothingKernel}.cpp),
 * rho-weighted gradient reconstruction (UF/VectorFiniteVolumeField.cpp:28-50), momentum-weighted face velocity,
 * variable-coefficient pressure equation.  Single rank.  Set gamma (cells and faces), then initialize.
 * fields: u p gradP gamma gradGamma rho mu beta sg fst kappa gammaTilde gradGammaTilde n gradRho
 * setup keys (before initialize): eps, kernelType (0 peskin, 1 pow6, 2 pow8), warmStart, contactAngle:<patch> (degrees)
 * stats: [itersGamma, itersU, itersP, relresGamma, relresU, relresP, maxDivergence, maxCourant] */
typedef struct phb_multiphase phb_multiphase;
int phb_mp_create(phb_mesh *m, double rho1, double rho2, double mu1, double mu2, double sigma, double gx, double gy,
                  double smoothingKernelRadius, phb_multiphase **out);
int phb_mp_destroy(phb_multiphase *mp);
phb_field *phb_mp_field(phb_multiphase *mp, const char *name);
phb_eqn *phb_mp_eqn(phb_multiphase *mp, const char *name);       /* gammaEqn uEqn pEqn */
phb_solver *phb_mp_solver(phb_multiphase *mp, const char *name);
int phb_mp_setup(phb_multiphase *mp, const char *key, double value);
int phb_mp_initialize(phb_multiphase *mp);
int phb_mp_step(phb_multiphase *mp, double dt, double stats[8]);

/* ------------------------------------------------------------ PISO time step
 * "phasePiso" of the north star.  The mounted snapshot has no PISO module any more
 * (SURVEY.md section 0); this driver is rebuilt from README.md:26-37, the commented
 * relax() body (UE/ScalarFiniteVolumeEquation.cpp:45-55) and the legacy case keys
 * (Examples/LidDrivenCavity/case/case.info:12-15).  Parity: self-consistency only. */
typedef struct phb_piso phb_piso;
int phb_piso_create(phb_mesh *m, double rho, double mu, phb_piso **out);
int phb_piso_destroy(phb_piso *s);
phb_field *phb_piso_field(phb_piso *s, const char *name);   /* u p pCorr gradP d */
phb_solver *phb_piso_solver(phb_piso *s, const char *name); /* uEqn pCorrEqn */
/* numInnerIterations numPressureCorrections momentumRelaxation pressureCorrectionRelaxation */
int phb_piso_setup(phb_piso *s, const char *key, double value);
int phb_piso_initialize(phb_piso *s);
/* stats: [itersU, itersPCorr, relresU, relresPCorr, maxMassImbalance, maxCourant] */
int phb_piso_step(phb_piso *s, double dt, double stats[6]);

#ifdef __cplusplus
}
#endif
#endif /* PHASE_B200_H */
