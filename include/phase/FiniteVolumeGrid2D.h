// FiniteVolumeGrid2D.h -- host handle of the mesh (reference:
// src/2D/Unstructured/FiniteVolumeGrid2D/FiniteVolumeGrid2D.{h,cpp},
// StructuredRectilinearGrid.{h,cpp}, Cell/Cell.h, Face/Face.h).  Connectivity,
// numbering and geometry are built by libphase_b200 (phb_mesh_*) with the
// reference's numbering; Cell / Face here are light value handles (id + cached
// geometry) so that solver-module code written against the reference --
// `for (const Cell &cell : *fluid_) u_(cell) += ...` -- compiles unchanged.
#ifndef PHASE_B200_FINITE_VOLUME_GRID_2D_H
#define PHASE_B200_FINITE_VOLUME_GRID_2D_H
#include <memory>
#include <string>
#include <vector>

#include "Communicator.h"
#include "Input.h"
#include "Vector2D.h"

class Cell {
public:
  Cell(Label id, Scalar volume, const Point2D &centroid) : id_(id), volume_(volume), centroid_(centroid) {}
  Label id() const { return id_; }
  Scalar volume() const { return volume_; }
  const Point2D &centroid() const { return centroid_; }

private:
  Label id_;
  Scalar volume_;
  Point2D centroid_;
};

class Face {
public:
  Face(Label id, const Point2D &centroid, const Vector2D &norm, Index lCell, Index rCell)
      : id_(id), centroid_(centroid), norm_(norm), lCell_(lCell), rCell_(rCell) {}
  Label id() const { return id_; }
  const Point2D &centroid() const { return centroid_; }
  const Vector2D &norm() const { return norm_; }  // area vector, outward from lCell
  bool isBoundary() const { return rCell_ < 0; }
  bool isInterior() const { return rCell_ >= 0; }
  Index lCellId() const { return lCell_; }
  Index rCellId() const { return rCell_; }

private:
  Label id_;
  Point2D centroid_;
  Vector2D norm_;
  Index lCell_, rCell_;
};

template <class T> class Group {
public:
  explicit Group(const std::string &name = "") : name_(name) {}
  const std::string &name() const { return name_; }
  void add(const T &item) { items_.push_back(std::cref(item)); }
  void add(const Group<T> &g) { items_.insert(items_.end(), g.items_.begin(), g.items_.end()); }
  void clear() { items_.clear(); }
  Size size() const { return items_.size(); }
  typename std::vector<Ref<const T>>::const_iterator begin() const { return items_.begin(); }
  typename std::vector<Ref<const T>>::const_iterator end() const { return items_.end(); }

private:
  std::string name_;
  std::vector<Ref<const T>> items_;
};
typedef Group<Cell> CellGroup;
typedef Group<Face> FaceGroup;

class FiniteVolumeGrid2D {
public:
  FiniteVolumeGrid2D(const std::shared_ptr<const Communicator> &comm, phb_mesh *mesh) : comm_(comm), m_(mesh) {
    phase::check(phb_mesh_finalize(m_), "FiniteVolumeGrid2D", "FiniteVolumeGrid2D");
    load();
  }
  // general unstructured input: nodes + CSR cell->node lists (FiniteVolumeGrid2D.cpp:12-35)
  FiniteVolumeGrid2D(const std::shared_ptr<const Communicator> &comm, const std::vector<Point2D> &nodes,
                     const std::vector<Label> &cptr, const std::vector<Label> &cind,
                     const std::vector<std::pair<std::string, std::vector<Label>>> &patchNodePairs = {})
      : comm_(comm) {
    std::vector<double> xy;
    for (const Point2D &p : nodes) { xy.push_back(p.x); xy.push_back(p.y); }
    std::vector<int> cp(cptr.begin(), cptr.end()), ci(cind.begin(), cind.end());
    phase::check(phb_mesh_create(comm->handle(), (int)nodes.size(), xy.data(), (int)cp.size() - 1, cp.data(),
                                 ci.data(), &m_), "FiniteVolumeGrid2D", "init");
    for (const auto &p : patchNodePairs) {
      std::vector<int> pr(p.second.begin(), p.second.end());
      phase::check(phb_mesh_add_patch_by_nodes(m_, p.first.c_str(), (int)pr.size() / 2, pr.data()),
                   "FiniteVolumeGrid2D", "createPatchByNodes");
      patchNames_.push_back(p.first);
    }
    phase::check(phb_mesh_finalize(m_), "FiniteVolumeGrid2D", "init");
    load();
  }
  FiniteVolumeGrid2D(const FiniteVolumeGrid2D &) = delete;
  virtual ~FiniteVolumeGrid2D() { phb_mesh_destroy(m_); }

  const Communicator &comm() const { return *comm_; }
  Size nCells() const { return cells_.size(); }
  Size nFaces() const { return faces_.size(); }
  const std::vector<Cell> &cells() const { return cells_; }
  const std::vector<Face> &faces() const { return faces_; }
  const CellGroup &localCells() const { return localCells_; }
  const std::vector<std::string> &patchNames() const { return patchNames_; }
  phb_mesh *handle() const { return m_; }
  std::vector<int> i32(const char *name) const {
    const long long n = phb_mesh_get_i32(m_, name, nullptr, 0);
    if (n < 0) throw Exception("FiniteVolumeGrid2D", "i32", phb_last_error());
    std::vector<int> v((size_t)n);
    phb_mesh_get_i32(m_, name, v.data(), n);
    return v;
  }
  std::vector<double> f64(const char *name) const {
    const long long n = phb_mesh_get_f64(m_, name, nullptr, 0);
    if (n < 0) throw Exception("FiniteVolumeGrid2D", "f64", phb_last_error());
    std::vector<double> v((size_t)n);
    phb_mesh_get_f64(m_, name, v.data(), n);
    return v;
  }
  // grid_->sendMessages(field): halo exchange of the cell values (FiniteVolumeGrid2D.tpp:3-49)
  template <class TField> void sendMessages(TField &field) const { field.sendMessages(); }

protected:
  void load() {
    const std::vector<double> vol = f64("vol"), cx = f64("cellCx"), cy = f64("cellCy");
    const std::vector<double> fx = f64("faceCx"), fy = f64("faceCy"), sx = f64("faceSx"), sy = f64("faceSy");
    const std::vector<int> fl = i32("faceL"), fr = i32("faceR"), owner = i32("owner");
    cells_.reserve(vol.size());
    for (size_t i = 0; i < vol.size(); ++i) cells_.push_back(Cell(i, vol[i], Point2D(cx[i], cy[i])));
    faces_.reserve(fl.size());
    for (size_t f = 0; f < fl.size(); ++f)
      faces_.push_back(Face(f, Point2D(fx[f], fy[f]), Vector2D(sx[f], sy[f]), fl[f], fr[f]));
    localCells_ = CellGroup("LocalCells");
    for (size_t i = 0; i < cells_.size(); ++i)
      if (owner[i] == comm_->rank()) localCells_.add(cells_[i]);
    static const char *std4[] = {"x-", "x+", "y-", "y+"};
    if (patchNames_.empty())
      for (const char *n : std4)
        if (phb_mesh_patch_id(m_, n) >= 0) patchNames_.push_back(n);
  }
  std::shared_ptr<const Communicator> comm_;
  phb_mesh *m_ = nullptr;
  std::vector<Cell> cells_;
  std::vector<Face> faces_;
  CellGroup localCells_;
  std::vector<std::string> patchNames_;
};

// StructuredRectilinearGrid(input): keys Grid.{width,height,nCellsX,nCellsY}
// (StructuredRectilinearGrid.cpp:3-32); patches x-, x+, y-, y+ (:175-194).
class StructuredRectilinearGrid : public FiniteVolumeGrid2D {
public:
  StructuredRectilinearGrid(const std::shared_ptr<const Communicator> &comm, Size nCellsX, Size nCellsY,
                            Scalar width, Scalar height)
      : FiniteVolumeGrid2D(comm, make(*comm, nCellsX, nCellsY, width, height)) {}
  StructuredRectilinearGrid(const std::shared_ptr<const Communicator> &comm, const Input &input)
      : StructuredRectilinearGrid(comm, input.caseInput().get<size_t>("Grid.nCellsX"),
                                  input.caseInput().get<size_t>("Grid.nCellsY"),
                                  input.caseInput().get<Scalar>("Grid.width") *
                                      input.caseInput().get<Scalar>("Grid.convertToMeters", 1.),
                                  input.caseInput().get<Scalar>("Grid.height") *
                                      input.caseInput().get<Scalar>("Grid.convertToMeters", 1.)) {}

private:
  static phb_mesh *make(const Communicator &comm, Size nx, Size ny, Scalar w, Scalar h) {
    phb_mesh *m = nullptr;
    if (comm.nProcs() == 1)
      phase::check(phb_mesh_create_rectilinear(comm.handle(), (int)nx, (int)ny, w, h, &m),
                   "StructuredRectilinearGrid", "init");
    else  // partitioned: y-strips, one per rank (the reference would call METIS here)
      phase::check(phb_mesh_create_rect_strip(comm.handle(), (int)nx, (int)ny, w, h, &m),
                   "StructuredRectilinearGrid", "init");
    return m;
  }
};
#endif
