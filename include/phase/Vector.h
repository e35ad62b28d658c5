// Vector.h -- dense right-hand-side vector, interface of src/Math/Vector.h:10-57.
#ifndef PHASE_B200_VECTOR_H
#define PHASE_B200_VECTOR_H
#include <algorithm>
#include <vector>

#include "Types.h"

class Vector {
public:
  Vector(Size size = 0, Scalar val = 0.) : data_(size, val) {}
  Scalar &operator()(Index i) { return data_[i]; }
  Scalar operator()(Index i) const { return data_[i]; }
  size_t size() const { return data_.size(); }
  void resize(Size size) { data_.resize(size); }
  void resize(Size size, Scalar val) { data_.resize(size, val); }
  void clear() { data_.clear(); }
  const std::vector<Scalar> &data() const { return data_; }
  std::vector<Scalar> &data() { return data_; }
  Vector &operator+=(const Vector &r) { for (size_t i = 0; i < data_.size(); ++i) data_[i] += r.data_[i]; return *this; }
  Vector &operator-=(const Vector &r) { for (size_t i = 0; i < data_.size(); ++i) data_[i] -= r.data_[i]; return *this; }
  Vector &operator+=(Scalar r) { for (Scalar &v : data_) v += r; return *this; }
  Vector &operator-=(Scalar r) { for (Scalar &v : data_) v -= r; return *this; }
  Vector &operator*=(Scalar r) { for (Scalar &v : data_) v *= r; return *this; }
  Vector &operator/=(Scalar r) { for (Scalar &v : data_) v /= r; return *this; }
  Vector operator-() const { Vector n(*this); for (Scalar &v : n.data_) v = -v; return n; }
  void zero() { std::fill(data_.begin(), data_.end(), 0.); }

private:
  std::vector<Scalar> data_;
};
inline Vector operator+(Vector l, const Vector &r) { return l += r; }
inline Vector operator-(Vector l, const Vector &r) { return l -= r; }
inline Vector operator*(Scalar l, Vector r) { return r *= l; }
inline Vector operator*(Vector l, Scalar r) { return l *= r; }
#endif
