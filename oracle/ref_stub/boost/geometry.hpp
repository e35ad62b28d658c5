// Stand-in for the part of Boost.Geometry the reference's 2D mesh classes call (TEST INFRASTRUCTURE: lets the
// reference's own grid / field / operator sources compile in place, see oracle/build.py).  Boost is not in this
// image.  Areas and centroids are the standard shoelace / Bashein-Detmer formulas Boost's cartesian strategies
// implement; they are OUR arithmetic, so cell volumes and centroids of the compiled reference are pinned to
// round-off only, everything else (numbering, links, operators, fields) is the reference's code.
#ifndef PHASE_ORACLE_BOOST_GEOMETRY_STUB
#define PHASE_ORACLE_BOOST_GEOMETRY_STUB
// the real headers pull in most of the standard library; the reference relies on that
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <deque>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <unordered_set>
#include <iterator>
#include <limits>
#include <memory>
#include <stdexcept>
#include <utility>
#include <vector>

#define BOOST_GEOMETRY_REGISTER_POINT_2D(Point, Coord, CS, X, Y)
#define BOOST_GEOMETRY_REGISTER_POINT_3D(Point, Coord, CS, X, Y, Z)

namespace boost { namespace geometry {
namespace cs { struct cartesian {}; }
namespace model {
template <class P, bool ClockWise = true, bool Closed = true> struct ring : public std::vector<P> {
  ring() {}
  template <class It> ring(It b, It e) : std::vector<P>(b, e) {}
};
template <class P> struct multi_point : public std::vector<P> {
  multi_point() {}
  template <class It> multi_point(It b, It e) : std::vector<P>(b, e) {}
};
template <class P> struct box {
  box() {}
  box(const P &lo, const P &hi) : lo_(lo), hi_(hi) {}
  const P &min_corner() const { return lo_; }
  const P &max_corner() const { return hi_; }
  P &min_corner() { return lo_; }
  P &max_corner() { return hi_; }
  P lo_, hi_;
};
template <class P, bool ClockWise = true, bool Closed = true> struct polygon {
  ring<P, ClockWise, Closed> &outer() { return outer_; }
  const ring<P, ClockWise, Closed> &outer() const { return outer_; }
  ring<P, ClockWise, Closed> outer_;
};
}  // namespace model

template <class R, class P> void append(R &r, const P &p) { r.push_back(p); }

template <class R> void unique(R &r) { r.erase(std::unique(r.begin(), r.end()), r.end()); }

// signed area of the ring as stored (counter-clockwise positive)
template <class R> double signed_area_ccw(const R &r) {
  double s = 0.;
  const std::size_t n = r.size();
  for (std::size_t i = 0; i + 1 < n; ++i) s += r[i].x * r[i + 1].y - r[i + 1].x * r[i].y;
  if (n > 1 && !(r.front() == r.back())) s += r[n - 1].x * r[0].y - r[0].x * r[n - 1].y;
  return 0.5 * s;
}
// ring<P, false, true>: counter-clockwise, closed
template <class P> void correct(model::ring<P, false, true> &r) {
  if (r.size() < 3) return;
  if (!(r.front() == r.back())) r.push_back(r.front());
  if (signed_area_ccw(r) < 0.) std::reverse(r.begin(), r.end());
}
template <class P> double area(const model::ring<P, false, true> &r) { return signed_area_ccw(r); }
template <class P> double perimeter(const model::ring<P, false, true> &r) {
  double s = 0.;
  for (std::size_t i = 0; i + 1 < r.size(); ++i) s += std::hypot(r[i + 1].x - r[i].x, r[i + 1].y - r[i].y);
  return s;
}
// Bashein-Detmer, coordinates relative to the first vertex
template <class P> void centroid(const model::ring<P, false, true> &r, P &c) {
  if (r.empty()) return;
  const double x0 = r[0].x, y0 = r[0].y;
  double a2 = 0., sx = 0., sy = 0.;
  for (std::size_t i = 0; i + 1 < r.size(); ++i) {
    const double x1 = r[i].x - x0, y1 = r[i].y - y0, x2 = r[i + 1].x - x0, y2 = r[i + 1].y - y0;
    const double ai = x1 * y2 - x2 * y1;
    a2 += ai; sx += ai * (x1 + x2); sy += ai * (y1 + y2);
  }
  if (a2 == 0.) { c = r[0]; return; }
  c.x = x0 + sx / (3. * a2);
  c.y = y0 + sy / (3. * a2);
}
template <class P> void envelope(const model::ring<P, false, true> &r, model::box<P> &b) {
  if (r.empty()) return;
  b.lo_ = b.hi_ = r[0];
  for (const P &p : r) {
    b.lo_.x = std::min(b.lo_.x, p.x); b.lo_.y = std::min(b.lo_.y, p.y);
    b.hi_.x = std::max(b.hi_.x, p.x); b.hi_.y = std::max(b.hi_.y, p.y);
  }
}
// -1 outside, 0 on the boundary, 1 inside
template <class P, class R> int point_in_ring(const P &p, const R &r) {
  bool in = false;
  const std::size_t n = r.size();
  for (std::size_t i = 0, j = n - 1; i < n; j = i++) {
    const double xi = r[i].x, yi = r[i].y, xj = r[j].x, yj = r[j].y;
    const double cross = (xj - xi) * (p.y - yi) - (yj - yi) * (p.x - xi);
    if (cross == 0. && std::min(xi, xj) <= p.x && p.x <= std::max(xi, xj) && std::min(yi, yj) <= p.y &&
        p.y <= std::max(yi, yj))
      return 0;
    if ((yi > p.y) != (yj > p.y) && p.x < (xj - xi) * (p.y - yi) / (yj - yi) + xi) in = !in;
  }
  return in ? 1 : -1;
}
template <class P> bool within(const P &p, const model::ring<P, false, true> &r) { return point_in_ring(p, r) > 0; }
template <class P> bool covered_by(const P &p, const model::ring<P, false, true> &r) { return point_in_ring(p, r) >= 0; }
template <class P> bool within(const P &p, const model::box<P> &b) {
  return p.x > b.lo_.x && p.x < b.hi_.x && p.y > b.lo_.y && p.y < b.hi_.y;
}
template <class P> bool covered_by(const P &p, const model::box<P> &b) {
  return p.x >= b.lo_.x && p.x <= b.hi_.x && p.y >= b.lo_.y && p.y <= b.hi_.y;
}
template <class R> bool is_valid(const R &r) { return r.size() >= 4 && signed_area_ccw(r) > 0.; }
template <class R> bool is_simple(const R &) { return true; }
template <class R> bool intersects(const R &a, const R &b) {
  for (const auto &p : a) if (point_in_ring(p, b) >= 0) return true;
  for (const auto &p : b) if (point_in_ring(p, a) >= 0) return true;
  return false;
}
template <class MP, class R> void convex_hull(const MP &pts, R &hull) {   // monotone chain, counter-clockwise, closed
  typedef typename MP::value_type P;
  std::vector<P> p(pts.begin(), pts.end());
  std::sort(p.begin(), p.end(), [](const P &a, const P &b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
  auto cr = [](const P &o, const P &a, const P &b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); };
  std::vector<P> h(2 * p.size() + 1);
  std::size_t k = 0;
  for (std::size_t i = 0; i < p.size(); ++i) { while (k >= 2 && cr(h[k - 2], h[k - 1], p[i]) <= 0) --k; h[k++] = p[i]; }
  for (std::size_t i = p.size() - 1, t = k + 1; i > 0; --i) { while (k >= t && cr(h[k - 2], h[k - 1], p[i - 1]) <= 0) --k; h[k++] = p[i - 1]; }
  hull.assign(h.begin(), h.begin() + k);
}
template <class R, class Out> void intersection(const R &, const R &, Out &) {
  throw std::logic_error("boost::geometry::intersection is not available in the oracle's Boost stand-in");
}
template <class R, class Out> void difference(const R &, const R &, Out &) {
  throw std::logic_error("boost::geometry::difference is not available in the oracle's Boost stand-in");
}

// ---- spatial index: a linear scan with the rtree's query interface
namespace index {
template <std::size_t A, std::size_t B> struct quadratic {};
namespace detail {
template <class F> struct Pred { F f; };   // f(point, item) -> keep?
template <class P> struct Nearest { P pt; std::size_t k; };
}  // namespace detail
template <class Geometry> struct WithinQ { Geometry g; };
template <class Geometry> struct CoveredQ { Geometry g; };
template <class F> struct SatisfiesQ { F f; };
template <class A, class B> struct AndQ { A a; B b; };
template <class Geometry> WithinQ<Geometry> within(const Geometry &g) { return WithinQ<Geometry>{g}; }
template <class Geometry> CoveredQ<Geometry> covered_by(const Geometry &g) { return CoveredQ<Geometry>{g}; }
template <class F> SatisfiesQ<F> satisfies(const F &f) { return SatisfiesQ<F>{f}; }
template <class P> detail::Nearest<P> nearest(const P &p, std::size_t k) { return detail::Nearest<P>{p, k}; }
template <class G, class B> AndQ<WithinQ<G>, B> operator&&(const WithinQ<G> &a, const B &b) { return {a, b}; }
template <class G, class B> AndQ<CoveredQ<G>, B> operator&&(const CoveredQ<G> &a, const B &b) { return {a, b}; }

// a query's results travel with its begin iterator (qbegin / qend are evaluated in unspecified order when both
// are arguments of one call); the end iterator is a sentinel
template <class Value> class query_iterator {
public:
  typedef std::forward_iterator_tag iterator_category;
  typedef Value value_type;
  typedef std::ptrdiff_t difference_type;
  typedef const Value *pointer;
  typedef const Value &reference;
  query_iterator() : i_(0) {}
  explicit query_iterator(std::shared_ptr<std::vector<Value>> r) : r_(std::move(r)), i_(0) {}
  reference operator*() const { return (*r_)[i_]; }
  pointer operator->() const { return &(*r_)[i_]; }
  query_iterator &operator++() { ++i_; return *this; }
  query_iterator operator++(int) { query_iterator t(*this); ++i_; return t; }
  bool at_end() const { return !r_ || i_ >= r_->size(); }
  bool operator==(const query_iterator &o) const {
    if (at_end() || o.at_end()) return at_end() && o.at_end();
    return r_ == o.r_ && i_ == o.i_;
  }
  bool operator!=(const query_iterator &o) const { return !(*this == o); }
private:
  std::shared_ptr<std::vector<Value>> r_;
  std::size_t i_;
};

template <class Value, class Params, class Getter, class Equal> class rtree {
public:
  typedef query_iterator<Value> const_query_iterator;
  void insert(const Value &v) { items_.push_back(v); }
  template <class It> void insert(It b, It e) { for (; b != e; ++b) items_.push_back(Value(*b)); }
  template <class T> std::size_t remove(const T &v) {
    Equal eq;
    for (auto it = items_.begin(); it != items_.end(); ++it)
      if (eq(*it, v)) { items_.erase(it); return 1; }
    return 0;
  }
  template <class It> std::size_t remove(It b, It e) { std::size_t n = 0; for (; b != e; ++b) n += remove(*b); return n; }
  void clear() { items_.clear(); }
  std::size_t size() const { return items_.size(); }
  template <class Q> const_query_iterator qbegin(const Q &q) const {
    auto out = std::make_shared<std::vector<Value>>();
    run(q, *out);
    return const_query_iterator(out);
  }
  const_query_iterator qend() const { return const_query_iterator(); }

private:
  template <class P, class Item> bool keep(const WithinQ<P> &q, const Item &it) const { return boost::geometry::within(Getter()(it), q.g); }
  template <class P, class Item> bool keep(const CoveredQ<P> &q, const Item &it) const { return boost::geometry::covered_by(Getter()(it), q.g); }
  template <class F, class Item> bool keep(const SatisfiesQ<F> &q, const Item &it) const { return q.f(it); }
  template <class A, class B, class Item> bool keep(const AndQ<A, B> &q, const Item &it) const { return keep(q.a, it) && keep(q.b, it); }
  template <class Q> void run(const Q &q, std::vector<Value> &out) const {
    for (const Value &v : items_)
      if (keep(q, v.get())) out.push_back(v);
  }
  template <class P> void run(const detail::Nearest<P> &q, std::vector<Value> &out) const {
    std::vector<std::pair<double, std::size_t>> d;
    for (std::size_t i = 0; i < items_.size(); ++i) {
      const auto c = Getter()(items_[i].get());
      d.push_back({(c.x - q.pt.x) * (c.x - q.pt.x) + (c.y - q.pt.y) * (c.y - q.pt.y), i});
    }
    const std::size_t k = std::min(q.k, d.size());
    std::partial_sort(d.begin(), d.begin() + k, d.end());
    for (std::size_t i = 0; i < k; ++i) out.push_back(items_[d[i].second]);
  }
  std::vector<Value> items_;
};
}  // namespace index
}}  // namespace boost::geometry
#endif
