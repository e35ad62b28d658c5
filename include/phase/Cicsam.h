// Cicsam.h -- namespace cicsam of the reference (src/2D/Unstructured/FiniteVolume/
// Discretization/Cicsam.{h,cpp}) on the mirrored API.  The face weights beta_f, which the
// reference passes around as std::vector<Scalar>, live in the FACE values of a scalar field so
// that they stay on the device between faceInterpolationWeights() and div().
#ifndef PHASE_B200_CICSAM_H
#define PHASE_B200_CICSAM_H
#include "FiniteVolumeEquation.h"

namespace cicsam {
// UD/Cicsam.cpp:19-66
inline void faceInterpolationWeights(const VectorFiniteVolumeField &u, const ScalarFiniteVolumeField &gamma,
                                     const VectorFiniteVolumeField &gradGamma, Scalar timeStep,
                                     ScalarFiniteVolumeField &beta) {
  phase::check(phb_cicsam_weights(u.handle(), gamma.handle(), gradGamma.handle(), timeStep, beta.handle()), "cicsam",
               "faceInterpolationWeights");
  beta.markDeviceDirty();
}
// UD/Cicsam.cpp:69-87
inline void computeMomentumFlux(Scalar rho1, Scalar rho2, const VectorFiniteVolumeField &u,
                                const ScalarFiniteVolumeField &gamma, const ScalarFiniteVolumeField &beta,
                                VectorFiniteVolumeField &rhoU) {
  phase::check(phb_cicsam_momentum_flux(rho1, rho2, u.handle(), gamma.handle(), beta.handle(), rhoU.handle()), "cicsam",
               "computeMomentumFlux");
  rhoU.markDeviceDirty();
}
// UD/Cicsam.cpp:89-144:  gammaEqn_ = (fv::ddt(gamma, dt) + cicsam::div(u, gamma, beta, 0.5) == 0.)
inline FiniteVolumeEquation<Scalar> div(const VectorFiniteVolumeField &u, ScalarFiniteVolumeField &gamma,
                                        const ScalarFiniteVolumeField &beta, Scalar theta) {
  FiniteVolumeEquation<Scalar> eqn(gamma);
  eqn.terms().push_back({phase::Term::CICSAM_DIV, 1., gamma.handle(), u.handle(), beta.handle(), 0., 0., theta, nullptr, {}});
  return eqn;
}
}  // namespace cicsam
#endif
