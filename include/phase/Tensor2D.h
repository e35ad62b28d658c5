// Tensor2D.h -- 2 x 2 tensor with the reference's member names (2D/Geometry/Tensor2D.h): what
// FiniteVolumeEquation<Vector2D>::add(cell, nb, Tensor2D) takes (UE/VectorFiniteVolumeEquation.cpp:48-66).
#ifndef PHASE_B200_TENSOR_2D_H
#define PHASE_B200_TENSOR_2D_H
#include "Vector2D.h"

class Tensor2D {
public:
  Tensor2D(Scalar xx = 0., Scalar xy = 0., Scalar yx = 0., Scalar yy = 0.) : xx(xx), xy(xy), yx(yx), yy(yy) {}
  Tensor2D &operator*=(Scalar a) { xx *= a; xy *= a; yx *= a; yy *= a; return *this; }
  Tensor2D &operator+=(const Tensor2D &o) { xx += o.xx; xy += o.xy; yx += o.yx; yy += o.yy; return *this; }
  Scalar xx, xy, yx, yy;
};
inline Tensor2D outer(const Vector2D &u, const Vector2D &v) { return Tensor2D(u.x * v.x, u.x * v.y, u.y * v.x, u.y * v.y); }
inline Tensor2D operator*(Scalar a, Tensor2D t) { return t *= a; }
inline Tensor2D operator*(Tensor2D t, Scalar a) { return t *= a; }
inline Vector2D dot(const Tensor2D &t, const Vector2D &u) { return Vector2D(t.xx * u.x + t.xy * u.y, t.yx * u.x + t.yy * u.y); }
#endif
