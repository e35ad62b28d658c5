// FiniteVolumeField.h -- FiniteVolumeField<T> (T = Scalar, Vector2D): host values
// with a device mirror (reference: src/2D/Unstructured/FiniteVolume/Field/
// FiniteVolumeField.{h,tpp}, ScalarFiniteVolumeField.cpp, VectorFiniteVolumeField.cpp,
// ScalarGradient.{h,cpp}).  The device copy is authoritative while kernels work on
// the field; host accessors synchronise lazily, so reference-style per-cell loops
// keep working (at the price of a transfer) while field-level calls stay on the GPU.
#ifndef PHASE_B200_FINITE_VOLUME_FIELD_H
#define PHASE_B200_FINITE_VOLUME_FIELD_H
#include <map>

#include "FiniteVolumeGrid2D.h"

template <class T> struct FieldTraits;
template <> struct FieldTraits<Scalar> {
  enum { nComp = 1 };
  static Scalar get(const double *v, size_t n, size_t i) { (void)n; return v[i]; }
  static void put(double *v, size_t n, size_t i, Scalar x) { (void)n; v[i] = x; }
  static void split(Scalar x, double &a, double &b) { a = x; b = 0.; }
  static Scalar parse(const std::string &s) { return std::stod(s); }
};
template <> struct FieldTraits<Vector2D> {
  enum { nComp = 2 };
  static Vector2D get(const double *v, size_t n, size_t i) { return Vector2D(v[i], v[n + i]); }
  static void put(double *v, size_t n, size_t i, const Vector2D &x) { v[i] = x.x; v[n + i] = x.y; }
  static void split(const Vector2D &x, double &a, double &b) { a = x.x; b = x.y; }
  static Vector2D parse(const std::string &s) { return Vector2D(s); }
};

template <class T> class FiniteVolumeField {
public:
  enum BoundaryType { FIXED, NORMAL_GRADIENT, SYMMETRY, OUTFLOW, PARTIAL_SLIP };
  enum InterpolationType { VOLUME, DISTANCE };

  FiniteVolumeField(const std::shared_ptr<const FiniteVolumeGrid2D> &grid, const std::string &name,
                    const T &val = T())
      : grid_(grid), name_(name), cells_(grid->nCells(), val), faces_(grid->nFaces(), val) {
    phase::check(phb_field_create(grid->handle(), FieldTraits<T>::nComp, name.c_str(), &f_), "FiniteVolumeField",
                 "FiniteVolumeField");
    hostDirty_ = true;
  }
  // boundary types + reference values from boundaries.info (FiniteVolumeField.tpp:425-478)
  FiniteVolumeField(const Input &input, const std::shared_ptr<const FiniteVolumeGrid2D> &grid,
                    const std::string &name, const T &val = T())
      : FiniteVolumeField(grid, name, val) {
    const auto &b = input.boundaryInput();
    for (const std::string &patch : grid->patchNames()) {
      std::string type = b.get<std::string>("Boundaries." + name + ".*.type", "");
      std::string value = b.get<std::string>("Boundaries." + name + ".*.value", "");
      const std::string t2 = b.get<std::string>("Boundaries." + name + "." + patch + ".type", "");
      const std::string v2 = b.get<std::string>("Boundaries." + name + "." + patch + ".value", "");
      if (!t2.empty()) type = t2;
      if (!v2.empty()) value = v2;
      if (type.empty()) continue;  // unlisted patches default to NORMAL_GRADIENT (:104-109)
      setBoundary(patch, parseType(type), value.empty() ? T() : FieldTraits<T>::parse(value));
    }
  }
  FiniteVolumeField(const FiniteVolumeField &) = delete;
  virtual ~FiniteVolumeField() { phb_field_destroy(f_); }

  const std::string &name() const { return name_; }
  const std::shared_ptr<const FiniteVolumeGrid2D> &grid() const { return grid_; }
  const CellGroup &cells() const { return grid_->localCells(); }

  void setBoundary(const std::string &patch, BoundaryType type, const T &ref) {
    if (type == OUTFLOW || type == PARTIAL_SLIP)
      throw Exception("FiniteVolumeField<T>", "setBoundary", "boundary type not supported by the b200 path.");
    double a, b2;
    FieldTraits<T>::split(ref, a, b2);
    toDevice();
    phase::check(phb_field_set_bc(f_, patch.c_str(), (int)type, a, b2), "FiniteVolumeField<T>", "setBoundary");
    bTypes_[patch] = type;
    deviceDirty_ = true;  // the patch faces took the reference value on the device
  }
  BoundaryType boundaryType(const std::string &patch) const {
    auto it = bTypes_.find(patch);
    return it == bTypes_.end() ? NORMAL_GRADIENT : it->second;
  }

  //- host access (synchronises lazily)
  T &operator()(const Cell &cell) { toHost(); hostDirty_ = true; return cells_[cell.id()]; }
  const T &operator()(const Cell &cell) const { toHost(); return cells_[cell.id()]; }
  T &operator()(const Face &face) { toHost(); hostDirty_ = true; return faces_[face.id()]; }
  const T &operator()(const Face &face) const { toHost(); return faces_[face.id()]; }
  T &operator[](Label id) { toHost(); hostDirty_ = true; return cells_[id]; }
  const T &operator[](Label id) const { toHost(); return cells_[id]; }
  void fill(const T &val) {
    double a, b;
    FieldTraits<T>::split(val, a, b);
    phase::check(phb_field_fill(f_, a, b), "FiniteVolumeField<T>", "fill");
    hostDirty_ = false; deviceDirty_ = true;
  }

  //- device-side field operations
  void interpolateFaces(InterpolationType type = DISTANCE) {
    if (type != DISTANCE) throw Exception("FiniteVolumeField<T>", "interpolateFaces", "only DISTANCE weights.");
    toDevice();
    phase::check(phb_field_interpolate_faces(f_), "FiniteVolumeField<T>", "interpolateFaces");
    deviceDirty_ = true;
  }
  void setBoundaryFaces() {
    toDevice();
    phase::check(phb_field_set_boundary_faces(f_), "FiniteVolumeField<T>", "setBoundaryFaces");
    deviceDirty_ = true;
  }
  // deep copy of cells + faces into the single history level (FiniteVolumeField.tpp:208-227)
  FiniteVolumeField<T> &savePreviousTimeStep(Scalar timeStep, int nPreviousFields) {
    (void)timeStep; (void)nPreviousFields;
    toDevice();
    phase::check(phb_field_save_previous(f_), "FiniteVolumeField<T>", "savePreviousTimeStep");
    return *this;
  }
  void sendMessages() {
    toDevice();
    phase::check(phb_field_send_messages(f_), "FiniteVolumeField<T>", "sendMessages");
    deviceDirty_ = true;
  }

  //- synchronisation
  void toDevice() const {
    if (!hostDirty_) return;
    const size_t N = cells_.size(), F = faces_.size();
    std::vector<double> c(FieldTraits<T>::nComp * N), f(FieldTraits<T>::nComp * F);
    for (size_t i = 0; i < N; ++i) FieldTraits<T>::put(c.data(), N, i, cells_[i]);
    for (size_t i = 0; i < F; ++i) FieldTraits<T>::put(f.data(), F, i, faces_[i]);
    phase::check(phb_field_set(f_, "cells", c.data(), (long long)c.size()), "FiniteVolumeField<T>", "toDevice");
    phase::check(phb_field_set(f_, "faces", f.data(), (long long)f.size()), "FiniteVolumeField<T>", "toDevice");
    hostDirty_ = false;
  }
  void toHost() const {
    if (!deviceDirty_) return;
    const size_t N = cells_.size(), F = faces_.size();
    std::vector<double> c(FieldTraits<T>::nComp * N), f(FieldTraits<T>::nComp * F);
    phase::check(phb_field_get(f_, "cells", c.data(), (long long)c.size()), "FiniteVolumeField<T>", "toHost");
    phase::check(phb_field_get(f_, "faces", f.data(), (long long)f.size()), "FiniteVolumeField<T>", "toHost");
    for (size_t i = 0; i < N; ++i) cells_[i] = FieldTraits<T>::get(c.data(), N, i);
    for (size_t i = 0; i < F; ++i) faces_[i] = FieldTraits<T>::get(f.data(), F, i);
    deviceDirty_ = false;
  }
  void markDeviceDirty() { deviceDirty_ = true; }
  phb_field *handle() const { toDevice(); return f_; }

  static BoundaryType parseType(const std::string &s) {
    if (s == "fixed") return FIXED;
    if (s == "normal_gradient") return NORMAL_GRADIENT;
    if (s == "symmetry") return SYMMETRY;
    if (s == "outflow") return OUTFLOW;
    if (s == "partial_slip") return PARTIAL_SLIP;
    throw Exception("FiniteVolumeField<T>", "setBoundaryTypes", "invalid boundary type \"" + s + "\".");
  }

protected:
  std::shared_ptr<const FiniteVolumeGrid2D> grid_;
  std::string name_;
  mutable std::vector<T> cells_, faces_;
  phb_field *f_ = nullptr;
  mutable bool hostDirty_ = false, deviceDirty_ = false;
  std::map<std::string, BoundaryType> bTypes_;
};

typedef FiniteVolumeField<Scalar> ScalarFiniteVolumeField;
typedef FiniteVolumeField<Vector2D> VectorFiniteVolumeField;

// ScalarGradient: face gradient + FACE_TO_CELL reconstruction (ScalarGradient.cpp:34-74)
class ScalarGradient : public VectorFiniteVolumeField {
public:
  enum Method { FACE_TO_CELL, GREEN_GAUSS_CELL, GREEN_GAUSS_NODE };
  explicit ScalarGradient(const ScalarFiniteVolumeField &phi)
      : VectorFiniteVolumeField(phi.grid(), "grad" + phi.name(), Vector2D(0., 0.)), phi_(phi) {}
  void compute(const CellGroup &, Method method = FACE_TO_CELL) { compute(method); }
  void compute(Method method = FACE_TO_CELL) {
    if (method != FACE_TO_CELL)
      throw Exception("ScalarGradient", "compute", "only FACE_TO_CELL is available on the b200 path.");
    toDevice();
    phase::check(phb_field_gradient(phi_.handle(), f_), "ScalarGradient", "compute");
    deviceDirty_ = true;
  }

private:
  const ScalarFiniteVolumeField &phi_;
};
#endif
