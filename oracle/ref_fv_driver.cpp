// ref_fv_driver.cpp -- C shim over the REFERENCE'S OWN finite-volume sources (TEST INFRASTRUCTURE ONLY).
//
// oracle/build.py compiles the reference's grid / field / equation / operator / FractionalStep translation units
// where they lie under /root/reference/src (nothing is copied) over the stand-in headers in oracle/ref_stub
// (Boost.Geometry, ptree, MPI with one rank, METIS, CGNS are not in this image) and links them with this file into
// oracle/_ref/libphase_ref_fv.so.  What runs behind these entry points is therefore the reference's code:
//   FiniteVolumeGrid2D::init / StructuredRectilinearGrid   (face numbering, links, patches)
//   FiniteVolumeField, ScalarGradient, IndexMap             (BCs, interpolation, gradient)
//   fv::ddt / div / laplacian, src::div / src               (UD/*.h, UD/*.cpp)
//   FiniteVolumeEquation<T>::solve                          (hand-off to the SparseMatrixSolver seam)
//   FractionalStep::solve                                   (US/FractionalStep.cpp)
// Ours: this shim, the SparseMatrixSolverFactory below (the reference's factory only knows Eigen / Trilinos
// backends, which are absent) that installs a RECORDING backend -- the position the B200 backend plugs into --
// and throwing stubs for the CGNS restart reader.  tests/ use it to pin oracle/phase_oracle.c and to generate
// tests/golden/ref_*.npz.
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "Math/SparseMatrixSolverFactory.h"
#include "System/CgnsFile.h"
#include "FiniteVolumeGrid2D/StructuredRectilinearGrid.h"
#include "Solvers/FractionalStep.h"
#include "Solvers/FractionalStepMultiphase.h"
#include "FiniteVolume/Discretization/Source.h"
#include "FiniteVolume/Discretization/TimeDerivative.h"

// ---- CGNS restart reader: never reached (the oracle does not restart)
CgnsFile::CgnsFile(const std::string &, Mode) { throw Exception("CgnsFile", "CgnsFile", "not available in the oracle build"); }
void CgnsFile::close() {}
CgnsFile::~CgnsFile() {}
std::vector<std::pair<std::string, std::string>> CgnsFile::readDescriptorNodes(int, int, int) const {
  throw Exception("CgnsFile", "readDescriptorNodes", "not available in the oracle build");
}
template <>
CgnsFile::Field<double> CgnsFile::readField<double>(int, int, int, int, int, const std::string &) const {
  throw Exception("CgnsFile", "readField", "not available in the oracle build");
}
template <>
CgnsFile::Field<int> CgnsFile::readField<int>(int, int, int, int, int, const std::string &) const {
  throw Exception("CgnsFile", "readField", "not available in the oracle build");
}
CgnsFile::Solution CgnsFile::readLastFlowSolution(int, int) const {
  throw Exception("CgnsFile", "readLastFlowSolution", "not available in the oracle build");
}

// ---- METIS hook: the partition vector is an input (oracle/ref_stub/metis.h)
#include <metis.h>
namespace { std::vector<int> g_part; }
void phase_metis_set_partition(const int *part, int n) { g_part.assign(part, part + n); }
int METIS_PartMeshDual(idx_t *ne, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, idx_t *, real_t *, idx_t *, idx_t *objval,
                       idx_t *epart, idx_t *) {
  if ((int)g_part.size() != *ne) return 0;
  std::copy(g_part.begin(), g_part.end(), epart);
  if (objval) *objval = 0;
  return METIS_OK;
}

// ---- the recording backend
typedef int (*rfv_solve_cb)(int n, const int *rowPtr, const int *colInd, const double *vals, const double *b, double *x,
                            void *user);
namespace {
rfv_solve_cb g_cb = nullptr;
void *g_user = nullptr;
std::string g_err;

class RecordingSolver : public SparseMatrixSolver {
public:
  Type type() const override { return EIGEN; }
  void setRank(int rank) override { n_ = rank; }
  void setRank(int rowRank, int) override { n_ = rowRank; }
  void set(const CoefficientList &) override { throw Exception("RecordingSolver", "set", "CSR hand-off only"); }
  void set(const std::vector<SparseEntry> &) override { throw Exception("RecordingSolver", "set", "CSR hand-off only"); }
  void set(const std::vector<Index> &rowPtr, const std::vector<Index> &colInds,
           const std::vector<Scalar> &vals) override {
    rowPtr_ = rowPtr; colInd_ = colInds; vals_ = vals;
  }
  void setGuess(const Vector &) override {}
  void setRhs(const Vector &rhs) override { b_ = rhs.data(); }
  Scalar solve() override {
    x_.assign(n_, 0.);
    if (!g_cb) throw Exception("RecordingSolver", "solve", "no solve callback installed");
    iters_ = g_cb(n_, rowPtr_.data(), colInd_.data(), vals_.data(), b_.data(), x_.data(), g_user);
    return 0.;
  }
  Scalar solve(const Vector &) override { return solve(); }
  Scalar x(Index idx) const override { return x_[idx]; }
  void setup(const boost::property_tree::ptree &) override {}
  int nIters() const override { return iters_; }
  Scalar error() const override { return 0.; }
  bool supportsMPI() const override { return false; }
  void printStatus(const std::string &) const override {}
  int n_ = 0, iters_ = 0;
  std::vector<Index> rowPtr_, colInd_;
  std::vector<Scalar> vals_, b_, x_;
};
}  // namespace

std::shared_ptr<SparseMatrixSolver> SparseMatrixSolverFactory::create(Type, const Communicator &) const {
  return std::make_shared<RecordingSolver>();
}
std::shared_ptr<SparseMatrixSolver> SparseMatrixSolverFactory::create(const std::string &, const Communicator &) const {
  return std::make_shared<RecordingSolver>();
}

namespace {
struct RefCase {
  Input input;
  explicit RefCase(const std::string &dir) : input(dir, dir + "/solution") { input.parseInputFile(); }
};
struct RefGrid {
  std::shared_ptr<FiniteVolumeGrid2D> g;
};
class OpenFracStep : public FractionalStep {
public:
  using FractionalStep::FractionalStep;
  using FractionalStep::gradP_;
  using FractionalStep::maxDivergenceError;
  using FractionalStep::p_;
  using FractionalStep::pEqn_;
  using FractionalStep::u_;
  using FractionalStep::uEqn_;
};
class OpenMultiphase : public FractionalStepMultiphase {
public:
  using FractionalStepMultiphase::FractionalStepMultiphase;
  using FractionalStepMultiphase::gammaEqn_;
  using FractionalStepMultiphase::maxDivergenceError;
  using FractionalStepMultiphase::pEqn_;
  using FractionalStepMultiphase::uEqn_;
};
struct RefFs {
  std::shared_ptr<OpenFracStep> fs;          // kind 0
  std::shared_ptr<OpenMultiphase> mp;        // kind 1
  std::shared_ptr<FiniteVolumeGrid2D> g;
  Solver &solver() { return fs ? static_cast<Solver &>(*fs) : static_cast<Solver &>(*mp); }
  FractionalStep &base() { return fs ? static_cast<FractionalStep &>(*fs) : static_cast<FractionalStep &>(*mp); }
};

template <class F> long guarded(F f) {
  try {
    return f();
  } catch (const std::exception &e) {
    g_err = e.what();
    return -1;
  } catch (...) {
    g_err = "unknown exception";
    return -1;
  }
}
}  // namespace

extern "C" {

const char *rfv_last_error() { return g_err.c_str(); }

void rfv_set_solver(rfv_solve_cb cb, void *user) { g_cb = cb; g_user = user; }

void *rfv_case_open(const char *dir) {
  RefCase *c = nullptr;
  if (guarded([&] { c = new RefCase(dir); return 0L; }) < 0) return nullptr;
  return c;
}
void rfv_case_close(void *c) { delete static_cast<RefCase *>(c); }

// StructuredRectilinearGrid(input): Grid.width/height/nCellsX/nCellsY of case.info
void *rfv_grid_rectilinear(void *caseHandle) {
  RefGrid *r = nullptr;
  if (guarded([&] {
        r = new RefGrid();
        r->g = std::make_shared<StructuredRectilinearGrid>(static_cast<RefCase *>(caseHandle)->input);
        return 0L;
      }) < 0) { delete r; return nullptr; }
  return r;
}
// FiniteVolumeGrid2D(nodes, cptr, cind, origin)
void *rfv_grid_create(int nNodes, const double *xy, int nCells, const int *cptr, const int *cind) {
  RefGrid *r = nullptr;
  if (guarded([&] {
        std::vector<Point2D> nodes;
        for (int i = 0; i < nNodes; ++i) nodes.emplace_back(xy[2 * i], xy[2 * i + 1]);
        std::vector<Label> cp(cptr, cptr + nCells + 1), ci(cind, cind + cptr[nCells]);
        r = new RefGrid();
        r->g = std::make_shared<FiniteVolumeGrid2D>(nodes, cp, ci, Point2D(0., 0.));
        return 0L;
      }) < 0) { delete r; return nullptr; }
  return r;
}
long rfv_grid_patch_by_nodes(void *grid, const char *name, int nNodes, const int *nodes) {
  return guarded([&] {
    std::vector<Label> n(nodes, nodes + nNodes);
    static_cast<RefGrid *>(grid)->g->createPatchByNodes(name, n);
    return 0L;
  });
}
void rfv_grid_close(void *g) { delete static_cast<RefGrid *>(g); }

// arrays in the oracle's naming (oracle/phase_oracle.c or_mesh_array); integers are written as doubles too when
// `asDouble`: the caller picks one buffer type per name.  out == NULL: count only.
long rfv_grid_get(void *grid, const char *name, int *iout, double *dout) {
  return guarded([&]() -> long {
    const FiniteVolumeGrid2D &g = *static_cast<RefGrid *>(grid)->g;
    const std::string nm(name);
    std::vector<long> iv;
    std::vector<double> dv;
    bool isInt = true;
    if (nm == "sizes") { iv = {(long)g.nNodes(), (long)g.nCells(), (long)g.nFaces(), (long)g.patches().size()}; }
    else if (nm == "nodeX" || nm == "nodeY") { isInt = false; for (const Node &n : g.nodes()) dv.push_back(nm == "nodeX" ? n.x : n.y); }
    else if (nm == "cptr") { iv.push_back(0); for (const Cell &c : g.cells()) iv.push_back(iv.back() + (long)c.nodes().size()); }
    else if (nm == "cind") { for (const Cell &c : g.cells()) for (const Node &n : c.nodes()) iv.push_back(n.id()); }
    else if (nm == "faceN1") { for (const Face &f : g.faces()) iv.push_back(f.lNode().id()); }
    else if (nm == "faceN2") { for (const Face &f : g.faces()) iv.push_back(f.rNode().id()); }
    else if (nm == "faceL") { for (const Face &f : g.faces()) iv.push_back(f.lCell().id()); }
    else if (nm == "faceR") { for (const Face &f : g.faces()) iv.push_back(f.isInterior() ? (long)f.rCell().id() : -1); }
    else if (nm == "faceCx" || nm == "faceCy") { isInt = false; for (const Face &f : g.faces()) dv.push_back(nm == "faceCx" ? f.centroid().x : f.centroid().y); }
    else if (nm == "faceNx" || nm == "faceNy") { isInt = false; for (const Face &f : g.faces()) dv.push_back(nm == "faceNx" ? f.norm().x : f.norm().y); }
    else if (nm == "faceW") { isInt = false; for (const Face &f : g.faces()) dv.push_back(f.isInterior() ? f.distanceWeight() : 1.); }
    else if (nm == "faceVW") { isInt = false; for (const Face &f : g.faces()) dv.push_back(f.isInterior() ? f.volumeWeight() : 1.); }
    else if (nm == "vol") { isInt = false; for (const Cell &c : g.cells()) dv.push_back(c.volume()); }
    else if (nm == "cellCx" || nm == "cellCy") { isInt = false; for (const Cell &c : g.cells()) dv.push_back(nm == "cellCx" ? c.centroid().x : c.centroid().y); }
    else if (nm == "ilPtr") { iv.push_back(0); for (const Cell &c : g.cells()) iv.push_back(iv.back() + (long)c.neighbours().size()); }
    else if (nm == "ilFace") { for (const Cell &c : g.cells()) for (const InteriorLink &l : c.neighbours()) iv.push_back(l.face().id()); }
    else if (nm == "ilCell") { for (const Cell &c : g.cells()) for (const InteriorLink &l : c.neighbours()) iv.push_back(l.cell().id()); }
    else if (nm == "ilRcx" || nm == "ilRcy") { isInt = false; for (const Cell &c : g.cells()) for (const InteriorLink &l : c.neighbours()) dv.push_back(nm == "ilRcx" ? l.rCellVec().x : l.rCellVec().y); }
    else if (nm == "ilSx" || nm == "ilSy") { isInt = false; for (const Cell &c : g.cells()) for (const InteriorLink &l : c.neighbours()) dv.push_back(nm == "ilSx" ? l.outwardNorm().x : l.outwardNorm().y); }
    else if (nm == "blPtr") { iv.push_back(0); for (const Cell &c : g.cells()) iv.push_back(iv.back() + (long)c.boundaries().size()); }
    else if (nm == "blFace") { for (const Cell &c : g.cells()) for (const BoundaryLink &l : c.boundaries()) iv.push_back(l.face().id()); }
    else if (nm == "blRfx" || nm == "blRfy") { isInt = false; for (const Cell &c : g.cells()) for (const BoundaryLink &l : c.boundaries()) dv.push_back(nm == "blRfx" ? l.rFaceVec().x : l.rFaceVec().y); }
    else if (nm == "blSx" || nm == "blSy") { isInt = false; for (const Cell &c : g.cells()) for (const BoundaryLink &l : c.boundaries()) dv.push_back(nm == "blSx" ? l.outwardNorm().x : l.outwardNorm().y); }
    else if (nm == "dlPtr") { iv.push_back(0); for (const Cell &c : g.cells()) iv.push_back(iv.back() + (long)c.diagonals().size()); }
    else if (nm == "dlCell") { for (const Cell &c : g.cells()) for (const CellLink &l : c.diagonals()) iv.push_back(l.cell().id()); }
    else if (nm.compare(0, 6, "patch:") == 0) {   // face ids of a named patch, in the patch's own order
      for (const Face &f : g.patch(nm.substr(6))) iv.push_back(f.id());
    } else throw Exception("rfv_grid_get", "name", "unknown array \"" + nm + "\"");
    if (isInt) {
      if (iout) for (size_t i = 0; i < iv.size(); ++i) iout[i] = (int)iv[i];
      if (dout) for (size_t i = 0; i < iv.size(); ++i) dout[i] = (double)iv[i];
      return (long)iv.size();
    }
    if (dout) std::copy(dv.begin(), dv.end(), dout);
    return (long)dv.size();
  });
}

// IndexMap(grid, nIndices): local / global index of every (cell, set)
long rfv_index_map(void *grid, int nIndices, int *local, int *global) {
  return guarded([&]() -> long {
    const FiniteVolumeGrid2D &g = *static_cast<RefGrid *>(grid)->g;
    IndexMap im(g, nIndices);
    long k = 0;
    for (int s = 0; s < nIndices; ++s)
      for (const Cell &c : g.cells()) {
        if (local) local[k] = (int)im.local(c, s);
        if (global) global[k] = (int)im.global(c, s);
        ++k;
      }
    return k;
  });
}

void *rfv_fs_create(void *caseHandle, void *grid) {
  RefFs *r = nullptr;
  if (guarded([&] {
        r = new RefFs();
        r->g = static_cast<RefGrid *>(grid)->g;
        r->fs = std::make_shared<OpenFracStep>(static_cast<RefCase *>(caseHandle)->input, r->g);
        r->fs->initialize();
        return 0L;
      }) < 0) { delete r; return nullptr; }
  return r;
}
void rfv_fs_close(void *fs) { delete static_cast<RefFs *>(fs); }

// FractionalStepMultiphase(input, grid) (US/FractionalStepMultiphase.cpp:11-50); NOT initialised: set gamma (and
// whatever else) through rfv_fs_any_field first, then rfv_fs_initialize
void *rfv_fsm_create(void *caseHandle, void *grid) {
  RefFs *r = nullptr;
  if (guarded([&] {
        r = new RefFs();
        r->g = static_cast<RefGrid *>(grid)->g;
        r->mp = std::make_shared<OpenMultiphase>(static_cast<RefCase *>(caseHandle)->input, r->g);
        return 0L;
      }) < 0) { delete r; return nullptr; }
  return r;
}
long rfv_fs_initialize(void *fs) {
  return guarded([&] {
    RefFs &R = *static_cast<RefFs *>(fs);
    if (R.fs) R.fs->initialize(); else R.mp->initialize();
    return 0L;
  });
}
long rfv_fs_step(void *fs, double dt) {
  return guarded([&] {
    RefFs &R = *static_cast<RefFs *>(fs);
    if (R.fs) R.fs->solve(dt); else R.mp->solve(dt);
    return 0L;
  });
}
double rfv_fs_max_divergence(void *fs) {
  RefFs &R = *static_cast<RefFs *>(fs);
  return R.fs ? R.fs->maxDivergenceError() : R.mp->maxDivergenceError();
}
double rfv_fs_max_courant(void *fs, double dt) { return static_cast<RefFs *>(fs)->base().maxCourantNumber(dt); }
double rfv_fs_max_time_step(void *fs, double maxCo, double prevDt) {
  RefFs &R = *static_cast<RefFs *>(fs);
  return R.fs ? R.fs->computeMaxTimeStep(maxCo, prevDt) : R.mp->computeMaxTimeStep(maxCo, prevDt);
}

// any registered field by its reference name (Solver::scalarField / vectorField: "gamma", "rho", "mu", "beta", "kappa",
// "gammaTilde", "u", "sg", "fst", "n", "rhoU", "gradgamma", ...): comp < 0 = scalar field, 0 / 1 = component of a
// vector field; faces != 0 = face values; set != 0 writes buf into the field.  Returns the length.
long rfv_fs_any_field(void *fsHandle, const char *name, int comp, int faces, double *buf, int set) {
  return guarded([&]() -> long {
    RefFs &R = *static_cast<RefFs *>(fsHandle);
    const FiniteVolumeGrid2D &g = *R.g;
    if (comp < 0) {
      std::shared_ptr<ScalarFiniteVolumeField> f = R.solver().scalarField(name);
      if (!f) throw Exception("rfv_fs_any_field", "name", std::string("no scalar field \"") + name + "\"");
      if (faces) { for (const Face &fc : g.faces()) { if (!buf) break; if (set) (*f)(fc) = buf[fc.id()]; else buf[fc.id()] = (*f)(fc); } return (long)g.nFaces(); }
      for (const Cell &c : g.cells()) { if (!buf) break; if (set) (*f)(c) = buf[c.id()]; else buf[c.id()] = (*f)(c); }
      return (long)g.nCells();
    }
    std::shared_ptr<VectorFiniteVolumeField> f = R.solver().vectorField(name);
    if (!f) throw Exception("rfv_fs_any_field", "name", std::string("no vector field \"") + name + "\"");
    if (faces) {
      for (const Face &fc : g.faces()) { if (!buf) break; Vector2D &v = (*f)(fc); if (set) (comp ? v.y : v.x) = buf[fc.id()]; else buf[fc.id()] = comp ? v.y : v.x; }
      return (long)g.nFaces();
    }
    for (const Cell &c : g.cells()) { if (!buf) break; Vector2D &v = (*f)(c); if (set) (comp ? v.y : v.x) = buf[c.id()]; else buf[c.id()] = comp ? v.y : v.x; }
    return (long)g.nCells();
  });
}

// field arrays in the oracle's naming: ux uy ufx ufy p pf gpx gpy gpfx gpfy ; set != 0 writes `buf` into the field
long rfv_fs_field(void *fsHandle, const char *name, double *buf, int set) {
  return guarded([&]() -> long {
    RefFs &R = *static_cast<RefFs *>(fsHandle);
    if (!R.fs) throw Exception("rfv_fs_field", "kind", "use rfv_fs_any_field on a multiphase solver");
    OpenFracStep &fs = *R.fs;
    const FiniteVolumeGrid2D &g = *R.g;
    const std::string nm(name);
    auto vecCells = [&](VectorFiniteVolumeField &f, int comp) -> long {
      for (const Cell &c : g.cells()) {
        if (!buf) continue;
        if (set) (comp ? f(c).y : f(c).x) = buf[c.id()];
        else buf[c.id()] = comp ? f(c).y : f(c).x;
      }
      return (long)g.nCells();
    };
    auto vecFaces = [&](VectorFiniteVolumeField &f, int comp) -> long {
      for (const Face &fc : g.faces()) {
        if (!buf) continue;
        if (set) (comp ? f(fc).y : f(fc).x) = buf[fc.id()];
        else buf[fc.id()] = comp ? f(fc).y : f(fc).x;
      }
      return (long)g.nFaces();
    };
    if (nm == "ux") return vecCells(fs.u_, 0);
    if (nm == "uy") return vecCells(fs.u_, 1);
    if (nm == "ufx") return vecFaces(fs.u_, 0);
    if (nm == "ufy") return vecFaces(fs.u_, 1);
    if (nm == "gpx") return vecCells(fs.gradP_, 0);
    if (nm == "gpy") return vecCells(fs.gradP_, 1);
    if (nm == "gpfx") return vecFaces(fs.gradP_, 0);
    if (nm == "gpfy") return vecFaces(fs.gradP_, 1);
    if (nm == "p") {
      for (const Cell &c : g.cells()) { if (!buf) continue; if (set) fs.p_(c) = buf[c.id()]; else buf[c.id()] = fs.p_(c); }
      return (long)g.nCells();
    }
    if (nm == "pf") {
      for (const Face &f : g.faces()) { if (!buf) continue; if (set) fs.p_(f) = buf[f.id()]; else buf[f.id()] = fs.p_(f); }
      return (long)g.nFaces();
    }
    throw Exception("rfv_fs_field", "name", "unknown field \"" + nm + "\"");
  });
}

// ---- FiniteVolumeGrid2D::partition + initCommBuffers + IndexMap on nRanks "MPI ranks" (threads, oracle/ref_mpi_threads.cpp)
// with the given partition vector: per rank the local grid's globalIds_, cellOwnership_, buffer / send cell groups
// (local cell ids, group order) and the IndexMap's local / global indices for nIndices sets.
namespace {
struct RankResult {
  std::vector<int> globalId, owner, bufPtr, bufCell, sendPtr, sendCell, local, global, faceL, faceR;
  std::string err;
};
struct PartJob {
  const RefCase *cs;
  int nIndices;
  std::vector<RankResult> res;
};
void part_worker(int rank, void *user) {
  PartJob &J = *static_cast<PartJob *>(user);
  RankResult &R = J.res[rank];
  try {
    StructuredRectilinearGrid g(J.cs->input);
    g.partition(J.cs->input);
    const int P = g.comm().nProcs();
    for (const Cell &c : g.cells()) { R.globalId.push_back((int)g.globalIds()[c.id()]); R.owner.push_back((int)g.cellOwnership()[c.id()]); }
    R.bufPtr.push_back(0); R.sendPtr.push_back(0);
    for (int q = 0; q < P; ++q) {
      for (const Cell &c : g.bufferGroups()[q]) R.bufCell.push_back((int)c.id());
      R.bufPtr.push_back((int)R.bufCell.size());
      for (const Cell &c : g.sendGroups()[q]) R.sendCell.push_back((int)c.id());
      R.sendPtr.push_back((int)R.sendCell.size());
    }
    for (const Face &f : g.faces()) { R.faceL.push_back((int)f.lCell().id()); R.faceR.push_back(f.isInterior() ? (int)f.rCell().id() : -1); }
    IndexMap im(g, J.nIndices);
    for (int s = 0; s < J.nIndices; ++s)
      for (const Cell &c : g.cells()) { R.local.push_back((int)im.local(c, s)); R.global.push_back((int)im.global(c, s)); }
  } catch (const std::exception &e) {
    R.err = e.what();
  }
}
PartJob *g_job = nullptr;
}  // namespace

long rfv_partition_run(void *caseHandle, int nRanks, const int *part, int nCells, int nIndices) {
  return guarded([&]() -> long {
    delete g_job;
    g_job = new PartJob();
    g_job->cs = static_cast<RefCase *>(caseHandle);
    g_job->nIndices = nIndices;
    g_job->res.resize(nRanks);
    phase_metis_set_partition(part, nCells);
    phase_mpi_run(nRanks, part_worker, g_job);
    for (auto &r : g_job->res)
      if (!r.err.empty()) throw Exception("rfv_partition_run", "rank", r.err);
    return 0L;
  });
}
// which: globalId owner bufPtr bufCell sendPtr sendCell local global faceL faceR
long rfv_partition_get(int rank, const char *which, int *out) {
  return guarded([&]() -> long {
    if (!g_job || rank < 0 || rank >= (int)g_job->res.size()) throw Exception("rfv_partition_get", "rank", "no such rank");
    const RankResult &R = g_job->res[rank];
    const std::string w(which);
    const std::vector<int> *v = w == "globalId" ? &R.globalId : w == "owner" ? &R.owner : w == "bufPtr" ? &R.bufPtr : w == "bufCell" ? &R.bufCell
                              : w == "sendPtr" ? &R.sendPtr : w == "sendCell" ? &R.sendCell : w == "local" ? &R.local : w == "global" ? &R.global
                              : w == "faceL" ? &R.faceL : w == "faceR" ? &R.faceR : nullptr;
    if (!v) throw Exception("rfv_partition_get", "which", "unknown array");
    if (out) std::copy(v->begin(), v->end(), out);
    return (long)v->size();
  });
}

// ---- operator probes on the fields of a FractionalStep (u_, p_, co_): the reference's own src:: / fv:: functions
// src::laplacian(Scalar gamma, p_) -> out[N]   (UD/Source.cpp:27-48)
long rfv_src_laplacian_scalar(void *fsHandle, double gamma, double *out) {
  return guarded([&]() -> long {
    OpenFracStep &fs = *static_cast<RefFs *>(fsHandle)->fs;
    Vector v = src::laplacian(gamma, fs.p_);
    std::copy(v.data().begin(), v.data().end(), out);
    return (long)v.size();
  });
}
// src::div(u_, cells) -> out[N]   (UD/Source.cpp:5-21)
long rfv_src_div_cells(void *fsHandle, int nCells, const int *cells, double *out) {
  return guarded([&]() -> long {
    RefFs &R = *static_cast<RefFs *>(fsHandle);
    CellGroup grp("probe");
    for (int i = 0; i < nCells; ++i) grp.add(R.g->cells()[cells[i]]);
    Vector v = src::div(R.fs->u_, grp);
    std::copy(v.data().begin(), v.data().end(), out);
    return (long)v.size();
  });
}
// fv::ddt(p_, dt, cells) after p_.savePreviousTimeStep -> diagonal coefficient and source per cell   (UD/TimeDerivative.h:50-62)
long rfv_ddt_cells(void *fsHandle, double dt, int nCells, const int *cells, double *diag, double *rhs) {
  return guarded([&]() -> long {
    RefFs &R = *static_cast<RefFs *>(fsHandle);
    CellGroup grp("probe");
    for (int i = 0; i < nCells; ++i) grp.add(R.g->cells()[cells[i]]);
    R.fs->p_.savePreviousTimeStep(dt, 1);
    FiniteVolumeEquation<Scalar> eqn = fv::ddt(R.fs->p_, dt, grp);
    const long n = (long)R.g->nCells();
    for (long i = 0; i < n; ++i) {
      diag[i] = 0.;
      for (Index k = eqn.rowPtr()[i]; k < eqn.rowPtr()[i + 1]; ++k)
        if (eqn.colInd()[k] == i) diag[i] += eqn.vals()[k];
      rhs[i] = eqn.b(i);
    }
    return n;
  });
}

// what FiniteVolumeEquation<T>::solve handed to the backend in the last solve of `which` ("uEqn" / "pEqn"):
// rowPtr (n+1), colInd / vals (rowPtr[n], -1 padded), b = -rhs_ (n).  NULL pointers: sizes only (n, nnz).
long rfv_fs_handoff(void *fsHandle, const char *which, int *rowPtr, int *colInd, double *vals, double *b, long *nnz) {
  return guarded([&]() -> long {
    RefFs &R = *static_cast<RefFs *>(fsHandle);
    const std::string w(which);
    std::shared_ptr<SparseMatrixSolver> sp;
    if (R.fs) sp = w == "uEqn" ? R.fs->uEqn_.sparseSolver() : R.fs->pEqn_.sparseSolver();
    else sp = w == "uEqn" ? R.mp->uEqn_.sparseSolver() : w == "pEqn" ? R.mp->pEqn_.sparseSolver() : R.mp->gammaEqn_.sparseSolver();
    RecordingSolver *s = static_cast<RecordingSolver *>(sp.get());
    if (nnz) *nnz = (long)s->colInd_.size();
    if (rowPtr) std::copy(s->rowPtr_.begin(), s->rowPtr_.end(), rowPtr);
    if (colInd) std::copy(s->colInd_.begin(), s->colInd_.end(), colInd);
    if (vals) std::copy(s->vals_.begin(), s->vals_.end(), vals);
    if (b) std::copy(s->b_.begin(), s->b_.end(), b);
    return (long)s->n_;
  });
}

}  // extern "C"
