// Stand-in for <boost/filesystem.hpp> over std::filesystem (TEST INFRASTRUCTURE): the restart code of the
// reference's Solver.cpp names path / directory_iterator / exists; the oracle never restarts.
#ifndef PHASE_ORACLE_BOOST_FILESYSTEM_STUB
#define PHASE_ORACLE_BOOST_FILESYSTEM_STUB
#include <cstdarg>
#include <filesystem>
#include <string>
namespace boost { namespace filesystem {
class path {
public:
  path() {}
  path(const std::string &s) : p_(s) {}
  path(const char *s) : p_(s) {}
  path(const std::filesystem::path &p) : p_(p) {}
  const std::string &string() const { s_ = p_.string(); return s_; }   // boost returns a reference
  path filename() const { return path(p_.filename()); }
  path parent_path() const { return path(p_.parent_path()); }
  path &operator/=(const path &o) { p_ /= o.p_; return *this; }
  const std::filesystem::path &std_path() const { return p_; }
private:
  std::filesystem::path p_;
  mutable std::string s_;
};
inline path operator/(const path &a, const path &b) { path r(a); r /= b; return r; }
inline bool exists(const path &p) { return std::filesystem::exists(p.std_path()); }
inline bool is_directory(const path &p) { return std::filesystem::is_directory(p.std_path()); }
inline bool create_directory(const path &p) { return std::filesystem::create_directory(p.std_path()); }
inline bool create_directories(const path &p) { return std::filesystem::create_directories(p.std_path()); }
class directory_entry {
public:
  directory_entry() {}
  explicit directory_entry(const std::filesystem::path &p) : p_(p) {}
  const filesystem::path &path() const { return p_; }
private:
  filesystem::path p_;
};
class directory_iterator {
public:
  directory_iterator() {}
  directory_iterator(const filesystem::path &p) : it_(p.std_path()) { sync(); }
  directory_iterator(const char *p) : it_(std::filesystem::path(p)) { sync(); }
  directory_iterator &operator++() { ++it_; sync(); return *this; }
  const directory_entry &operator*() const { return cur_; }
  const directory_entry *operator->() const { return &cur_; }
  bool operator!=(const directory_iterator &o) const { return it_ != o.it_; }
  bool operator==(const directory_iterator &o) const { return it_ == o.it_; }
private:
  void sync() { if (it_ != std::filesystem::directory_iterator()) cur_ = directory_entry(it_->path()); }
  std::filesystem::directory_iterator it_;
  directory_entry cur_;
};
}}
#endif
