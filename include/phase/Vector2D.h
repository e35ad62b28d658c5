// Vector2D.h -- the subset of src/2D/Geometry/Vector2D.h the hot path touches.
#ifndef PHASE_B200_VECTOR_2D_H
#define PHASE_B200_VECTOR_2D_H
#include <cmath>
#include <string>

#include "Types.h"

class Vector2D {
public:
  Vector2D(Scalar x = 0., Scalar y = 0.) : x(x), y(y) {}
  // "(x,y)" as written in the .info files (Vector2D.cpp:14-26)
  Vector2D(std::string s) {
    const size_t a = s.find_first_of("("), b = s.find_last_of(")");
    s = s.substr(a + 1, b - a - 1);
    const size_t c = s.find_first_of(", \t");
    x = std::stod(s.substr(0, c));
    y = std::stod(s.substr(s.find_first_not_of(", \t", c)));
  }
  Scalar magSqr() const { return x * x + y * y; }
  Scalar mag() const { return std::sqrt(x * x + y * y); }
  Vector2D abs() const { return Vector2D(std::abs(x), std::abs(y)); }
  Vector2D unitVec() const { return Vector2D(x / mag(), y / mag()); }
  Vector2D normalVec() const { return Vector2D(y, -x); }
  Vector2D tangentVec() const { return Vector2D(-y, x); }
  Vector2D &operator+=(const Vector2D &o) { x += o.x; y += o.y; return *this; }
  Vector2D &operator-=(const Vector2D &o) { x -= o.x; y -= o.y; return *this; }
  Vector2D &operator*=(Scalar s) { x *= s; y *= s; return *this; }
  Vector2D &operator/=(Scalar s) { x /= s; y /= s; return *this; }
  bool operator==(const Vector2D &o) const { return x == o.x && y == o.y; }
  Scalar x, y;
};
typedef Vector2D Point2D;
inline Vector2D operator+(Vector2D a, const Vector2D &b) { return a += b; }
inline Vector2D operator-(Vector2D a, const Vector2D &b) { return a -= b; }
inline Vector2D operator-(const Vector2D &a) { return Vector2D(-a.x, -a.y); }
inline Vector2D operator*(Vector2D a, Scalar s) { return a *= s; }
inline Vector2D operator*(Scalar s, Vector2D a) { return a *= s; }
inline Vector2D operator/(Vector2D a, Scalar s) { return a /= s; }
inline Scalar dot(const Vector2D &a, const Vector2D &b) { return a.x * b.x + a.y * b.y; }
inline Scalar cross(const Vector2D &a, const Vector2D &b) { return a.x * b.y - a.y * b.x; }
inline Vector2D pointwise(const Vector2D &a, const Vector2D &b) { return Vector2D(a.x * b.x, a.y * b.y); }
#endif
