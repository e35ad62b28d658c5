// FractionalStep.h -- the snapshot's solver module written against the mirrored
// API; the bodies of solveUEqn / solvePEqn / correctVelocity are the reference's
// statements verbatim in meaning and order (src/2D/Unstructured/Solvers/
// FractionalStep.cpp:25-135), which is the point: a module written for Phase runs
// on the b200 path without modification.  (The reference derives from Solver --
// field registry, ICs, restart -- which is outside the hot path; fields are plain
// members here.)
#ifndef PHASE_B200_FRACTIONAL_STEP_H
#define PHASE_B200_FRACTIONAL_STEP_H
#include <cmath>

#include "FiniteVolumeEquation.h"

class FractionalStep {
public:
  FractionalStep(const Input &input, const std::shared_ptr<const FiniteVolumeGrid2D> &grid)
      : grid_(grid), fluid_(std::make_shared<CellGroup>("fluid")), u_(input, grid, "u", Vector2D(0., 0.)),
        p_(input, grid, "p", 0.), gradP_(p_), uEqn_(input, u_, "uEqn"), pEqn_(input, p_, "pEqn") {
    fluid_->add(grid_->localCells());
    rho_ = input.caseInput().get<Scalar>("Properties.rho", 1);
    mu_ = input.caseInput().get<Scalar>("Properties.mu", 1);
    maxTimeStep_ = input.caseInput().get<Scalar>("Solver.timeStep", 1e300);
  }
  virtual ~FractionalStep() {}

  virtual void initialize() {
    u_.interpolateFaces();
    p_.setBoundaryFaces();
  }

  virtual Scalar solve(Scalar timeStep) {
    solveUEqn(timeStep);
    solvePEqn(timeStep);
    correctVelocity(timeStep);
    return 0;
  }

  VectorFiniteVolumeField &u() { return u_; }
  ScalarFiniteVolumeField &p() { return p_; }
  ScalarGradient &gradP() { return gradP_; }
  FiniteVolumeEquation<Vector2D> &uEqn() { return uEqn_; }
  FiniteVolumeEquation<Scalar> &pEqn() { return pEqn_; }

protected:
  virtual Scalar solveUEqn(Scalar timeStep) {
    u_.savePreviousTimeStep(timeStep, 1);

    uEqn_ = (fv::ddt(u_, timeStep) + fv::div(u_, u_, 0.) ==
             fv::laplacian(mu_ / rho_, u_, 0.5) - src::src(gradP_));

    Scalar error = uEqn_.solve();

    for (const Cell &cell : *fluid_)
      u_(cell) += timeStep * gradP_(cell);

    grid_->sendMessages(u_);
    u_.interpolateFaces();

    return error;
  }

  virtual Scalar solvePEqn(Scalar timeStep) {
    pEqn_ = (fv::laplacian(timeStep, p_) == src::div(u_));

    Scalar error = pEqn_.solve();
    grid_->sendMessages(p_);
    p_.setBoundaryFaces();

    //- Gradient
    gradP_.compute(*fluid_);

    return error;
  }

  virtual void correctVelocity(Scalar timeStep) {
    for (const Cell &cell : *fluid_)
      u_(cell) -= timeStep * gradP_(cell);

    grid_->sendMessages(u_);  //- Necessary

    for (const Face &face : grid_->faces())
      u_(face) -= timeStep * gradP_(face);
  }

  std::shared_ptr<const FiniteVolumeGrid2D> grid_;
  Scalar rho_, mu_, maxTimeStep_;
  std::shared_ptr<CellGroup> fluid_;
  VectorFiniteVolumeField u_;
  ScalarFiniteVolumeField p_;
  ScalarGradient gradP_;
  FiniteVolumeEquation<Vector2D> uEqn_;
  FiniteVolumeEquation<Scalar> pEqn_;
};
#endif
