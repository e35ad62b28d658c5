// FiniteVolumeEquation.h -- Seam 2: FiniteVolumeEquation<T> and the fv:: / src::
// free functions with the reference's names and argument meaning, so solver
// modules written against Phase compile unchanged:
//
//   uEqn_ = (fv::ddt(u_, dt) + fv::div(u_, u_, 0.) == fv::laplacian(mu_/rho_, u_, 0.5) - src::src(gradP_));
//   uEqn_.solve();
//
// Reference: src/2D/Unstructured/FiniteVolume/Equation/FiniteVolumeEquation.{h,tpp},
// ScalarFiniteVolumeEquation.cpp, VectorFiniteVolumeEquation.cpp, and
// Discretization/{TimeDerivative,Divergence,ExplicitDivergence,Laplacian,Source}.*
//
// B200 design: the operator functions do NOT build intermediate matrices (the
// reference allocates a 5-wide padded CSR per operator and rebuilds the whole
// matrix in every + / - / ==, M/CrsEquation.cpp:185-275).  Each returns a light
// equation carrying a term list; the assignment into the persistent member
// equation (uEqn_, pEqn_) zeroes the device equation once and runs one
// accumulating kernel per term on the canonical pattern (phb_assemble_*).  The
// per-entry add()/set() API of the reference is kept as a host fall-through
// (CrsEquation semantics) for callers that build arbitrary stencils.
#ifndef PHASE_B200_FINITE_VOLUME_EQUATION_H
#define PHASE_B200_FINITE_VOLUME_EQUATION_H
#include <algorithm>
#include <cctype>

#include <valarray>

#include "CrsEquation.h"
#include "FiniteVolumeField.h"
#include "SparseMatrixSolverFactory.h"
#include "Tensor2D.h"

namespace phase {
struct Term {
  enum Kind { DDT, DIV, DIVE, LAPLACIAN, SRC, SRC_DIV, CICSAM_DIV, SRC_LAPLACIAN, DDT_CELLS, SRC_DIV_CELLS } kind;
  double sign;
  phb_field *phi, *u, *aux;  // phi: transported/solved field, u: advecting field, aux: rho / gamma field
  double c0, c1, c2;         // rho or gamma constant, dt, theta
  phb_field *rowScale;       // rho * (sub-expression)
  std::vector<int> cells;    // cell-group overloads: ids of the group's cells
};
// right-hand-side vectors of src:: (evaluated on the device when they meet an equation)
struct Source {
  std::vector<Term> terms;
};
inline Source operator+(Source l, const Source &r) { l.terms.insert(l.terms.end(), r.terms.begin(), r.terms.end()); return l; }
inline Source operator-(Source l, const Source &r) {
  for (Term t : r.terms) { t.sign = -t.sign; l.terms.push_back(t); }
  return l;
}
inline Source operator-(Source s) { for (Term &t : s.terms) t.sign = -t.sign; return s; }
}  // namespace phase

template <class T> class FiniteVolumeEquation : public CrsEquation {
public:
  FiniteVolumeEquation(FiniteVolumeField<T> &field, const std::string &name = "", int nnz = 5)
      : name(name), field_(field), nnz_(nnz) {}
  FiniteVolumeEquation(FiniteVolumeField<T> &field, int nnz) : FiniteVolumeEquation(field, "", nnz) {}
  FiniteVolumeEquation(const Input &input, FiniteVolumeField<T> &field, const std::string &name, int nnz = 5)
      : FiniteVolumeEquation(field, name, nnz) {
    configureSparseSolver(input, field.grid()->comm());
  }
  FiniteVolumeEquation(const FiniteVolumeEquation<T> &o)
      : CrsEquation(o), name(o.name), field_(o.field_), nnz_(o.nnz_), terms_(o.terms_), hostUsed_(o.hostUsed_) {}
  ~FiniteVolumeEquation() { if (e_) phb_eqn_destroy(e_); }

  // assignment keeps the destination's solver (M/CrsEquation.cpp:16-26) and, for
  // operator expressions, runs the device assembly
  FiniteVolumeEquation<T> &operator=(const FiniteVolumeEquation<T> &rhs) {
    if (this == &rhs) return *this;
    if (rhs.hostUsed_) {
      if (!rhs.terms_.empty())
        throw Exception("FiniteVolumeEquation<T>", "operator=", "cannot mix per-entry add() with fv:: operators.");
      CrsEquation::operator=(static_cast<const CrsEquation &>(rhs));
      hostUsed_ = true; deviceReady_ = false;
      return *this;
    }
    terms_ = rhs.terms_;
    assembleOnDevice();
    return *this;
  }
  FiniteVolumeEquation<T> &operator=(const CrsEquation &rhs) {
    CrsEquation::operator=(rhs);
    hostUsed_ = true; deviceReady_ = false;
    return *this;
  }

  //- Add/set/get coefficients: host fall-through with the reference's semantics
  void set(const Cell &cell, const Cell &nb, Scalar val) { each(cell, nb, [&](Index r, Index c) { setCoeff(r, c, val); }); }
  void add(const Cell &cell, const Cell &nb, Scalar val) { each(cell, nb, [&](Index r, Index c) { addCoeff(r, c, val); }); }
  void addSource(const Cell &cell, Scalar val) { ensureHost(); for (int k = 0; k < nComp(); ++k) rhs_(row(cell, k)) += val; }
  void setSource(const Cell &cell, Scalar val) { ensureHost(); for (int k = 0; k < nComp(); ++k) rhs_(row(cell, k)) = val; }
  void addSource(const Cell &cell, const Vector2D &v) { ensureHost(); rhs_(row(cell, 0)) += v.x; rhs_(row(cell, 1)) += v.y; }
  void scale(const Cell &cell, Scalar val) { ensureHost(); for (int k = 0; k < nComp(); ++k) scaleRow(row(cell, k), val); }
  void remove(const Cell &cell) {
    ensureHost();
    for (int k = 0; k < nComp(); ++k) {
      const Index r = row(cell, k);
      std::fill(colInd_.begin() + rowPtr_[r], colInd_.begin() + rowPtr_[r + 1], -1);
      rhs_(r) = 0.;
    }
  }
  template <typename cell_iterator, typename coeff_iterator>
  void add(const Cell &cell, cell_iterator begin, cell_iterator end, coeff_iterator coeffs) {
    for (cell_iterator itr = begin; itr != end; ++itr, ++coeffs) add(cell, *itr, *coeffs);
  }
  // container overloads (UE/FiniteVolumeEquation.h:52-63)
  void add(const Cell &cell, const std::vector<Ref<const Cell>> &nbs, const std::vector<Scalar> &vals) {
    for (size_t i = 0; i < nbs.size(); ++i) add(cell, nbs[i].get(), vals[i]);
  }
  void add(const Cell &cell, const std::vector<Ref<const Cell>> &nbs, const std::valarray<Scalar> &vals) {
    for (size_t i = 0; i < nbs.size(); ++i) add(cell, nbs[i].get(), vals[i]);
  }
  // component-wise and tensor coefficients of a vector equation (UE/VectorFiniteVolumeEquation.cpp:38-66): the
  // off-diagonal tensor entries are only created when non-zero, "to avoid coupling whenever possible"
  void add(const Cell &cell, const Cell &nb, const Vector2D &val) {
    requireVector("add");
    ensureHost();
    addCoeff(row(cell, 0), col(nb, 0), val.x);
    addCoeff(row(cell, 1), col(nb, 1), val.y);
  }
  void add(const Cell &cell, const Cell &nb, const Tensor2D &val) {
    requireVector("add");
    ensureHost();
    addCoeff(row(cell, 0), col(nb, 0), val.xx);
    if (val.xy != 0.) addCoeff(row(cell, 0), col(nb, 1), val.xy);
    if (val.yx != 0.) addCoeff(row(cell, 1), col(nb, 0), val.yx);
    addCoeff(row(cell, 1), col(nb, 1), val.yy);
  }
  // get(cell, nb): the coefficient(s) of nb in the row(s) of cell (UE/VectorFiniteVolumeEquation.cpp:68-76)
  T get(const Cell &cell, const Cell &nb) {
    ensureHost();
    return getImpl(cell, nb, static_cast<T *>(nullptr));
  }

  // LinearAlgebra.<name>.lib selects the backend (FiniteVolumeEquation.tpp:47-62)
  void configureSparseSolver(const Input &input, const Communicator &comm) {
    std::string lib = input.caseInput().get<std::string>("LinearAlgebra." + name + ".lib");
    std::transform(lib.begin(), lib.end(), lib.begin(), [](unsigned char ch) { return (char)std::tolower(ch); });
    solver_ = SparseMatrixSolverFactory().create(lib, comm);
    if (comm.nProcs() > 1 && !solver_->supportsMPI())
      throw Exception("FiniteVolumeEquation<T>", "configureSparseSolver",
                      "equation \"" + name + "\", lib \"" + lib + "\" does not support multiple processes.");
    solver_->setup(input.caseInput().get_child("LinearAlgebra." + name));
    comm.printf("Initialized sparse matrix solver for equation \"%s\" using lib%s.\n", name.c_str(), lib.c_str());
  }

  // FiniteVolumeEquation<T>::solve (FiniteVolumeEquation.tpp:64-86)
  Scalar solve() {
    if (!solver_)
      throw Exception("FiniteVolumeEquation<T>", "solve",
                      "must allocate a SparseMatrixSolver object before attempting to solve.");
    if (deviceReady_) {
      B200SparseMatrixSolver *b = dynamic_cast<B200SparseMatrixSolver *>(solver_.get());
      if (!b) throw Exception("FiniteVolumeEquation<T>", "solve", "device equations need the b200 backend.");
      int iters = 0;
      double err = 0.;
      const int rc = phb_eqn_solve(e_, b->handle(), field_.handle(), warmStart ? 1 : 0, &iters, &err);
      b->setLastResult(iters, err);
      phase::check(rc, "FiniteVolumeEquation<T>", "solve");
      field_.markDeviceDirty();
    } else {
      solver_->setRank((int)getRank());
      solver_->set(rowPtr_, colInd_, vals_);
      solver_->setRhs(-rhs_);
      solver_->solve();
      mapFromSparseSolver();
    }
    solver_->printStatus("FiniteVolumeEquation " + name + ":");
    return solver_->error();
  }

  // relax(omega): a_PP /= omega; rhs_P -= (1 - omega) a_PP phi_P  (body removed from
  // the snapshot, recovered from ScalarFiniteVolumeEquation.cpp:45-55)
  void relax(Scalar relaxationFactor) {
    if (deviceReady_) {
      phase::check(phb_eqn_relax(e_, field_.handle(), relaxationFactor), "FiniteVolumeEquation<T>", "relax");
      return;
    }
    ensureHost();
    for (const Cell &cell : field_.cells())
      for (int k = 0; k < nComp(); ++k) {
        const Index r = row(cell, k);
        for (Index j = rowPtr_[r]; j < rowPtr_[r + 1]; ++j)
          if (colInd_[j] == col(cell, k)) {
            vals_[j] /= relaxationFactor;
            rhs_(r) -= (1. - relaxationFactor) * vals_[j] * component(field_(cell), k);
          }
      }
  }

  const FiniteVolumeField<T> &field() const { return field_; }
  // parity export in the reference's CSR layout (layout: see phb_eqn_export_csr)
  void exportReferenceLayout(int layout, std::vector<Index> &rowPtr, std::vector<Index> &colInd,
                             std::vector<Scalar> &vals, std::vector<Scalar> &rhs) const {
    if (!deviceReady_) { rowPtr = rowPtr_; colInd = colInd_; vals = vals_; rhs = rhs_.data(); return; }
    const long long nnz = phb_eqn_export_csr(e_, layout, nullptr, nullptr, nullptr, nullptr);
    if (nnz < 0) throw Exception("FiniteVolumeEquation<T>", "exportReferenceLayout", phb_last_error());
    const Size n = getRank();
    rowPtr.resize(n + 1); colInd.resize(nnz); vals.resize(nnz); rhs.resize(n);
    phb_eqn_export_csr(e_, layout, rowPtr.data(), colInd.data(), vals.data(), rhs.data());
  }

  std::string name;
  bool warmStart = false;  // the reference passes no guess; opt-in extension

  //- expression building (used by fv:: / src:: and the operators below)
  std::vector<phase::Term> &terms() { return terms_; }
  const std::vector<phase::Term> &terms() const { return terms_; }
  bool hostUsed() const { return hostUsed_; }

protected:
  void requireVector(const char *method) const {
    if (FieldTraits<T>::nComp != 2)
      throw Exception("FiniteVolumeEquation<T>", method, "Vector2D / Tensor2D coefficients need a vector equation.");
  }
  Scalar getImpl(const Cell &cell, const Cell &nb, Scalar *) { return coeff(row(cell, 0), col(nb, 0)); }
  Vector2D getImpl(const Cell &cell, const Cell &nb, Vector2D *) {
    return Vector2D(coeff(row(cell, 0), col(nb, 0)), coeff(row(cell, 1), col(nb, 1)));
  }
  int nComp() const { return FieldTraits<T>::nComp; }
  static Scalar component(const Scalar &v, int) { return v; }
  static Scalar component(const Vector2D &v, int k) { return k == 0 ? v.x : v.y; }
  Size nLocal() const { return field_.grid()->localCells().size(); }
  Size getRank() const { return nComp() * nLocal(); }
  void loadIndexMap() {
    if (!localRow_.empty()) return;
    localRow_ = field_.grid()->i32("localRow");
    globalRow_ = field_.grid()->i32("globalRow");
  }
  // IndexMap: local(cell,k) = k nLocal + local, global = offset + local (UE/IndexMap.cpp:29-36)
  Index row(const Cell &c, int k) { loadIndexMap(); return (Index)(k * nLocal()) + localRow_[c.id()]; }
  Index col(const Cell &c, int k) {
    loadIndexMap();
    if (k > 0 && field_.grid()->comm().nProcs() > 1)
      throw Exception("FiniteVolumeEquation<T>", "add", "per-entry vector equations are single-process only.");
    return (Index)(k * nLocal()) + globalRow_[c.id()];
  }
  template <class F> void each(const Cell &cell, const Cell &nb, F f) {
    ensureHost();
    for (int k = 0; k < nComp(); ++k) f(row(cell, k), col(nb, k));
  }
  void ensureHost() {
    if (!terms_.empty())
      throw Exception("FiniteVolumeEquation<T>", "add", "cannot mix per-entry add() with fv:: operators.");
    if (!hostUsed_) {
      static_cast<CrsEquation &>(*this) = CrsEquation(getRank(), (Size)nnz_);
      hostUsed_ = true; deviceReady_ = false;
    }
  }
  void mapFromSparseSolver() {
    for (const Cell &cell : field_.cells()) assign(field_(cell), cell);
  }
  void assign(Scalar &v, const Cell &c) { v = solver_->x(row(c, 0)); }
  void assign(Vector2D &v, const Cell &c) { v.x = solver_->x(row(c, 0)); v.y = solver_->x(row(c, 1)); }

  void assembleOnDevice() {
    if (!e_) phase::check(phb_eqn_create(field_.grid()->handle(), nComp(), &e_), "FiniteVolumeEquation<T>", "operator=");
    phase::check(phb_eqn_zero(e_), "FiniteVolumeEquation<T>", "operator=");
    // rho * (sub-expression): scaled terms first, one row scaling, then the rest
    phb_field *scale = nullptr;
    for (const phase::Term &t : terms_)
      if (t.rowScale) {
        if (scale && scale != t.rowScale)
          throw Exception("FiniteVolumeEquation<T>", "operator=", "only one row-scaling field per expression.");
        scale = t.rowScale;
      }
    for (int pass = 0; pass < 2; ++pass) {
      for (const phase::Term &t : terms_) {
        if ((pass == 0) != (t.rowScale != nullptr)) continue;
        int rc = 0;
        switch (t.kind) {
        case phase::Term::DDT: rc = phb_assemble_ddt(e_, t.phi, t.c0, t.aux, t.c1, t.sign); break;
        case phase::Term::DIV: rc = phb_assemble_div(e_, t.u, t.phi, t.c2, t.sign); break;
        case phase::Term::DIVE: rc = phb_assemble_dive(e_, t.u, t.phi, t.c2, t.sign); break;
        case phase::Term::LAPLACIAN: rc = phb_assemble_laplacian(e_, t.c0, t.aux, t.phi, t.c2, t.sign); break;
        case phase::Term::SRC: rc = phb_assemble_src(e_, t.phi, t.sign); break;
        case phase::Term::SRC_DIV: rc = phb_assemble_src_div(e_, t.u, t.sign); break;
        case phase::Term::CICSAM_DIV: rc = phb_assemble_cicsam_div(e_, t.u, t.phi, t.aux, t.c2, t.sign); break;
        case phase::Term::SRC_LAPLACIAN: rc = phb_assemble_src_laplacian(e_, t.c0, t.aux, t.phi, t.sign); break;
        case phase::Term::DDT_CELLS:
          rc = phb_assemble_ddt_cells(e_, t.phi, t.c1, t.sign, (int)t.cells.size(), t.cells.data());
          break;
        case phase::Term::SRC_DIV_CELLS:
          rc = phb_assemble_src_div_cells(e_, t.u, t.sign, (int)t.cells.size(), t.cells.data());
          break;
        }
        phase::check(rc, "FiniteVolumeEquation<T>", "operator=");
      }
      if (pass == 0 && scale) phase::check(phb_eqn_scale_rows(e_, scale), "FiniteVolumeEquation<T>", "operator=");
    }
    deviceReady_ = true; hostUsed_ = false;
  }

  FiniteVolumeField<T> &field_;
  int nnz_;
  std::vector<phase::Term> terms_;
  bool hostUsed_ = false, deviceReady_ = false;
  phb_eqn *e_ = nullptr;
  std::vector<int> localRow_, globalRow_;
};

// ---------------------------------------------------------------- operators
namespace phase {
template <class T> void append(FiniteVolumeEquation<T> &l, const FiniteVolumeEquation<T> &r, double sign) {
  if (l.hostUsed() || r.hostUsed()) {
    if (!l.terms().empty() || !r.terms().empty())
      throw Exception("FiniteVolumeEquation<T>", "operator", "cannot mix per-entry add() with fv:: operators.");
    if (sign > 0) static_cast<CrsEquation &>(l) += static_cast<const CrsEquation &>(r);
    else static_cast<CrsEquation &>(l) -= static_cast<const CrsEquation &>(r);
    return;
  }
  for (Term t : r.terms()) { t.sign *= sign; l.terms().push_back(t); }
}
template <class T> void append(FiniteVolumeEquation<T> &l, const Source &s, double sign) {
  if (l.hostUsed()) throw Exception("FiniteVolumeEquation<T>", "operator", "src:: terms need a device equation.");
  for (Term t : s.terms) { t.sign *= sign; l.terms().push_back(t); }
}
}  // namespace phase

template <class T> FiniteVolumeEquation<T> operator+(FiniteVolumeEquation<T> l, const FiniteVolumeEquation<T> &r) { phase::append(l, r, +1.); return l; }
template <class T> FiniteVolumeEquation<T> operator-(FiniteVolumeEquation<T> l, const FiniteVolumeEquation<T> &r) { phase::append(l, r, -1.); return l; }
template <class T> FiniteVolumeEquation<T> operator==(FiniteVolumeEquation<T> l, const FiniteVolumeEquation<T> &r) { phase::append(l, r, -1.); return l; }
// eqn +/- Vector touches rhs_ only (M/CrsEquation.cpp:277-285); `==` is `-=` (:305-311)
template <class T> FiniteVolumeEquation<T> operator+(FiniteVolumeEquation<T> l, const phase::Source &s) { phase::append(l, s, +1.); return l; }
template <class T> FiniteVolumeEquation<T> operator-(FiniteVolumeEquation<T> l, const phase::Source &s) { phase::append(l, s, -1.); return l; }
template <class T> FiniteVolumeEquation<T> operator==(FiniteVolumeEquation<T> l, const phase::Source &s) { phase::append(l, s, -1.); return l; }
template <class T> FiniteVolumeEquation<T> operator==(FiniteVolumeEquation<T> l, Scalar rhs) {
  if (rhs != 0.) throw Exception("FiniteVolumeEquation<T>", "operator==", "only `== 0.` is supported for device equations.");
  return l;
}
// rho * eqn: row scaling of coefficients and rhs (VectorFiniteVolumeEquation.cpp:163-170)
template <class T> FiniteVolumeEquation<T> operator*(const ScalarFiniteVolumeField &rho, FiniteVolumeEquation<T> r) {
  if (r.hostUsed()) throw Exception("FiniteVolumeEquation<T>", "operator*", "row scaling needs a device equation.");
  for (phase::Term &t : r.terms()) t.rowScale = rho.handle();
  return r;
}

// ------------------------------------------------------------- fv:: and src::
namespace fv {
template <typename T> FiniteVolumeEquation<T> ddt(Scalar rho, FiniteVolumeField<T> &field, Scalar timeStep) {
  FiniteVolumeEquation<T> eqn(field);
  eqn.terms().push_back({phase::Term::DDT, 1., field.handle(), nullptr, nullptr, rho, timeStep, 0., nullptr, {}});
  return eqn;
}
template <typename T>
FiniteVolumeEquation<T> ddt(const ScalarFiniteVolumeField &rho, FiniteVolumeField<T> &field, Scalar timeStep) {
  FiniteVolumeEquation<T> eqn(field);
  eqn.terms().push_back({phase::Term::DDT, 1., field.handle(), nullptr, rho.handle(), 1., timeStep, 0., nullptr, {}});
  return eqn;
}
template <typename T> FiniteVolumeEquation<T> ddt(FiniteVolumeField<T> &field, Scalar timeStep) {
  return ddt(1., field, timeStep);
}
// the cells of a group only (UD/TimeDerivative.h:50-62; the immersed-boundary modules use it)
template <typename T> FiniteVolumeEquation<T> ddt(FiniteVolumeField<T> &field, Scalar timeStep, const CellGroup &cells) {
  FiniteVolumeEquation<T> eqn(field);
  phase::Term t = {phase::Term::DDT_CELLS, 1., field.handle(), nullptr, nullptr, 1., timeStep, 0., nullptr, {}};
  for (const Cell &c : cells) t.cells.push_back((int)c.id());
  eqn.terms().push_back(t);
  return eqn;
}
template <typename T>
FiniteVolumeEquation<T> div(const VectorFiniteVolumeField &u, FiniteVolumeField<T> &phi, Scalar theta = 1.) {
  FiniteVolumeEquation<T> eqn(phi);
  eqn.terms().push_back({phase::Term::DIV, 1., phi.handle(), u.handle(), nullptr, 0., 0., theta, nullptr, {}});
  return eqn;
}
template <typename T>
FiniteVolumeEquation<T> dive(const VectorFiniteVolumeField &u, FiniteVolumeField<T> &phi, Scalar theta) {
  FiniteVolumeEquation<T> eqn(phi);
  eqn.terms().push_back({phase::Term::DIVE, 1., phi.handle(), u.handle(), nullptr, 0., 0., theta, nullptr, {}});
  return eqn;
}
template <class T> FiniteVolumeEquation<T> laplacian(Scalar gamma, FiniteVolumeField<T> &phi, Scalar theta) {
  FiniteVolumeEquation<T> eqn(phi);
  eqn.terms().push_back({phase::Term::LAPLACIAN, 1., phi.handle(), nullptr, nullptr, gamma, 0., theta, nullptr, {}});
  return eqn;
}
template <class T> FiniteVolumeEquation<T> laplacian(Scalar gamma, FiniteVolumeField<T> &phi) {
  return laplacian(gamma, phi, -1.);  // steady overload: no old-time terms
}
template <class T>
FiniteVolumeEquation<T> laplacian(const ScalarFiniteVolumeField &gamma, FiniteVolumeField<T> &phi, Scalar theta) {
  FiniteVolumeEquation<T> eqn(phi);
  eqn.terms().push_back({phase::Term::LAPLACIAN, 1., phi.handle(), nullptr, gamma.handle(), 0., 0., theta, nullptr, {}});
  return eqn;
}
template <class T> FiniteVolumeEquation<T> laplacian(const ScalarFiniteVolumeField &gamma, FiniteVolumeField<T> &phi) {
  return laplacian(gamma, phi, -1.);
}
}  // namespace fv

namespace src {
inline phase::Source div(const VectorFiniteVolumeField &field) {
  return {{{phase::Term::SRC_DIV, 1., nullptr, field.handle(), nullptr, 0., 0., 0., nullptr, {}}}};
}
// src::div(field, cells) (UD/Source.cpp:5-21)
inline phase::Source div(const VectorFiniteVolumeField &field, const CellGroup &cells) {
  phase::Source s = {{{phase::Term::SRC_DIV_CELLS, 1., nullptr, field.handle(), nullptr, 0., 0., 0., nullptr, {}}}};
  for (const Cell &c : cells) s.terms[0].cells.push_back((int)c.id());
  return s;
}
// src::laplacian (UD/Source.cpp:27-75): sum over the links of gamma g_f (phi_nb - phi_P).  The reference's field
// overload (:50-75) indexes a one-component index map out of bounds; its evident intent (the same scalar sum with
// gamma taken at the faces) is what exists here.
inline phase::Source laplacian(Scalar gamma, const ScalarFiniteVolumeField &phi) {
  return {{{phase::Term::SRC_LAPLACIAN, 1., phi.handle(), nullptr, nullptr, gamma, 0., 0., nullptr, {}}}};
}
inline phase::Source laplacian(const ScalarFiniteVolumeField &gamma, const ScalarFiniteVolumeField &phi) {
  return {{{phase::Term::SRC_LAPLACIAN, 1., phi.handle(), nullptr, gamma.handle(), 0., 0., 0., nullptr, {}}}};
}
inline phase::Source src(const ScalarFiniteVolumeField &field) {
  return {{{phase::Term::SRC, 1., field.handle(), nullptr, nullptr, 0., 0., 0., nullptr, {}}}};
}
inline phase::Source src(const VectorFiniteVolumeField &field) {
  return {{{phase::Term::SRC, 1., field.handle(), nullptr, nullptr, 0., 0., 0., nullptr, {}}}};
}
}  // namespace src

namespace fv {
// README-era spelling: `== fv::laplacian(mu, u) - fv::grad(p)`: the pressure gradient as an
// explicit source, i.e. src::src of the reconstructed cell gradient (README.md:26-37)
inline phase::Source grad(ScalarGradient &gradP) {
  gradP.compute();
  return src::src(static_cast<const VectorFiniteVolumeField &>(gradP));
}
}  // namespace fv
#endif
