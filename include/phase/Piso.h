// Piso.h -- "phasePiso": the README-era SIMPLE/PISO solver module (README.md:26-37) on the
// mirrored API.  The mounted snapshot ships no such module any more, so
// this class reads the LEGACY case keys of Examples/LidDrivenCavity/case/case.info:9-29
// (Solver.numInnerIterations / numPressureCorrections / momentumRelaxation /
// pressureCorrectionRelaxation, LinearAlgebra.uEqn / pCorrEqn) and drives the device-resident
// implementation (phb_piso_*), whose momentum and pressure-correction equations are exactly
//   uEqn_     = (fv::ddt(rho,u,dt) + fv::div(rho*u,u) == fv::laplacian(mu,u) - fv::grad(p)); relax; solve
//   pCorrEqn_ = (fv::laplacian(rho*d, pCorr) == m); solve
#ifndef PHASE_B200_PISO_H
#define PHASE_B200_PISO_H
#include "FiniteVolumeGrid2D.h"
#include "Input.h"

class Piso {
public:
  Piso(const Input &input, const std::shared_ptr<const FiniteVolumeGrid2D> &grid) : grid_(grid) {
    const auto &ci = input.caseInput();
    phase::check(phb_piso_create(grid->handle(), ci.get<Scalar>("Properties.rho", 1.), ci.get<Scalar>("Properties.mu", 1.), &s_),
                 "Piso", "Piso");
    for (const char *k : {"numInnerIterations", "numPressureCorrections", "momentumRelaxation", "pressureCorrectionRelaxation"}) {
      const boost::property_tree::ptree *c = ci.get_child_optional(std::string("Solver.") + k);
      if (c) phase::check(phb_piso_setup(s_, k, std::stod(c->data())), "Piso", "Piso");
    }
    // boundary conditions of u and p (boundaries.info), same parsing rules as FiniteVolumeField
    const auto &b = input.boundaryInput();
    for (const char *fname : {"u", "p"}) {
      phb_field *f = phb_piso_field(s_, fname);
      for (const std::string &patch : grid->patchNames()) {
        std::string type = b.get<std::string>(std::string("Boundaries.") + fname + ".*.type", "");
        std::string value = b.get<std::string>(std::string("Boundaries.") + fname + ".*.value", "");
        const std::string t2 = b.get<std::string>(std::string("Boundaries.") + fname + "." + patch + ".type", "");
        const std::string v2 = b.get<std::string>(std::string("Boundaries.") + fname + "." + patch + ".value", "");
        if (!t2.empty()) type = t2;
        if (!v2.empty()) value = v2;
        if (type.empty()) continue;
        const int t = type == "fixed" ? PHB_FIXED : type == "normal_gradient" ? PHB_NORMAL_GRADIENT
                      : type == "symmetry" ? PHB_SYMMETRY : -1;
        if (t < 0) throw Exception("Piso", "Piso", "invalid boundary type \"" + type + "\".");
        Vector2D v(0., 0.);
        if (!value.empty()) v = value.find('(') != std::string::npos ? Vector2D(value) : Vector2D(std::stod(value), 0.);
        phase::check(phb_field_set_bc(f, patch.c_str(), t, v.x, v.y), "Piso", "Piso");
      }
    }
    // LinearAlgebra.<eqn>: every key except `lib` is forwarded to the backend
    for (const char *eq : {"uEqn", "pCorrEqn"}) {
      const boost::property_tree::ptree *la = ci.get_child_optional(std::string("LinearAlgebra.") + eq);
      if (!la) continue;
      std::string lib = la->get<std::string>("lib", "b200");
      if (lib != "b200")
        throw Exception("SparseMatrixSolverFactory", "create", "bad solver type \"" + lib + "\".");
      for (const auto &kv : *la)
        if (kv.first != "lib")
          phase::check(phb_solver_setup(phb_piso_solver(s_, eq), kv.first.c_str(), kv.second.data().c_str()), "Piso", "Piso");
    }
  }
  Piso(const Piso &) = delete;
  ~Piso() { phb_piso_destroy(s_); }
  void initialize() { phase::check(phb_piso_initialize(s_), "Piso", "initialize"); }
  // returns the maximum mass imbalance after the step
  Scalar solve(Scalar timeStep) {
    double st[6];
    phase::check(phb_piso_step(s_, timeStep, st), "Piso", "solve");
    grid_->comm().printf("Max mass imbalance = %.4e, max CFL = %.4lf, iterations u/pCorr = %d/%d\n", st[4], st[5], (int)st[0], (int)st[1]);
    return st[4];
  }
  std::vector<double> field(const char *name, const char *part = "cells") const {
    phb_field *f = phb_piso_field(s_, name);
    if (!f) throw Exception("Piso", "field", std::string("no field ") + name);
    const bool vec = std::string(name) == "u" || std::string(name) == "gradP";
    const size_t n = (vec ? 2 : 1) * (std::string(part) == "cells" ? grid_->nCells() : grid_->nFaces());
    std::vector<double> v(n);
    phase::check(phb_field_get(f, part, v.data(), (long long)n), "Piso", "field");
    return v;
  }

private:
  std::shared_ptr<const FiniteVolumeGrid2D> grid_;
  phb_piso *s_ = nullptr;
};
#endif
