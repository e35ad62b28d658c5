// Stand-in for <boost/property_tree/ptree.hpp> (Boost is not installed here): TEST INFRASTRUCTURE used to
// compile the reference's own sources IN PLACE for oracle/_ref.  Implements the calls those sources make:
// get<T>(path[, default]), get_child, get_child_optional (.get(), bool), put, data, iteration over children.
#ifndef PHASE_ORACLE_PTREE_STUB
#define PHASE_ORACLE_PTREE_STUB
#include <algorithm>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>
namespace boost {
template <class T> class optional;
template <class T> class optional<T &> {
public:
  optional() : p_(nullptr) {}
  optional(T &v) : p_(&v) {}
  explicit operator bool() const { return p_ != nullptr; }
  bool operator!() const { return p_ == nullptr; }
  T &get() const { return *p_; }
  T &operator*() const { return *p_; }
  T *operator->() const { return p_; }
private:
  T *p_;
};
namespace property_tree {
class ptree_error : public std::runtime_error { public: explicit ptree_error(const std::string &w) : std::runtime_error(w) {} };
class ptree_bad_path : public ptree_error { public: explicit ptree_bad_path(const std::string &w) : ptree_error(w) {} };
class ptree_bad_data : public ptree_error { public: explicit ptree_bad_data(const std::string &w) : ptree_error(w) {} };
class ptree {
public:
  typedef std::string key_type;
  typedef std::string data_type;
  typedef std::pair<const std::string, ptree> value_type;
  typedef std::vector<std::pair<std::string, ptree>> Children;
  typedef Children::const_iterator const_iterator;
  typedef Children::iterator iterator;
  ptree() {}
  explicit ptree(const std::string &d) : data_(d) {}
  const std::string &data() const { return data_; }
  std::string &data() { return data_; }
  const_iterator begin() const { return kids_.begin(); }
  const_iterator end() const { return kids_.end(); }
  iterator begin() { return kids_.begin(); }
  iterator end() { return kids_.end(); }
  bool empty() const { return kids_.empty(); }
  std::size_t size() const { return kids_.size(); }
  std::size_t count(const std::string &key) const {
    std::size_t n = 0;
    for (const auto &k : kids_) n += k.first == key;
    return n;
  }
  const ptree *walk(const std::string &path) const {
    const ptree *t = this;
    std::size_t pos = 0;
    if (path.empty()) return t;
    for (;;) {
      const std::size_t dot = path.find('.', pos);
      const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
      const ptree *next = nullptr;
      for (const auto &k : t->kids_)
        if (k.first == key) { next = &k.second; break; }
      if (!next) return nullptr;
      t = next;
      if (dot == std::string::npos) return t;
      pos = dot + 1;
    }
  }
  const ptree &get_child(const std::string &path) const {
    const ptree *t = walk(path);
    if (!t) throw ptree_bad_path("No such node (" + path + ")");
    return *t;
  }
  optional<const ptree &> get_child_optional(const std::string &path) const {
    const ptree *t = walk(path);
    return t ? optional<const ptree &>(*t) : optional<const ptree &>();
  }
  template <class T> T get_value() const { return convert<T>(data_); }
  template <class T> T get(const std::string &path) const { return convert<T>(get_child(path).data_); }
  template <class T> T get(const std::string &path, const T &def) const {   // default when absent OR untranslatable
    const ptree *t = walk(path);
    if (!t) return def;
    try { return convert<T>(t->data_); } catch (const ptree_bad_data &) { return def; }
  }
  std::string get(const std::string &path, const char *def) const { return get<std::string>(path, std::string(def)); }
  ptree &put(const std::string &path, const std::string &value) {
    ptree *t = this;
    std::size_t pos = 0;
    for (;;) {
      const std::size_t dot = path.find('.', pos);
      const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
      ptree *next = nullptr;
      for (auto &k : t->kids_)
        if (k.first == key) { next = &k.second; break; }
      if (!next) { t->kids_.push_back(std::make_pair(key, ptree())); next = &t->kids_.back().second; }
      t = next;
      if (dot == std::string::npos) break;
      pos = dot + 1;
    }
    t->data_ = value;
    return *t;
  }
  ptree &put(const std::string &path, const char *value) { return put(path, std::string(value)); }
  template <class T> ptree &put(const std::string &path, const T &value) {
    std::ostringstream os;
    os.precision(17);
    os << value;
    return put(path, os.str());
  }
  ptree &add_child(const std::string &key, const ptree &child) { kids_.push_back(std::make_pair(key, child)); return kids_.back().second; }
  Children &children() { return kids_; }
private:
  template <class T> static T convert(const std::string &s) {
    std::istringstream is(s);
    T v;
    is >> std::boolalpha >> v;
    if (is.fail()) {   // bool written as 0/1, on/off
      std::istringstream is2(s);
      is2 >> v;
      if (is2.fail()) throw ptree_bad_data("conversion of data \"" + s + "\" failed");
    }
    return v;
  }
  std::string data_;
  Children kids_;
};
template <> inline std::string ptree::convert<std::string>(const std::string &s) { return s; }
}  // namespace property_tree
}  // namespace boost
#endif
