// Stand-in for the few Boost.StringAlgo calls of the reference (TEST INFRASTRUCTURE, see boost/geometry.hpp).
#ifndef PHASE_ORACLE_BOOST_STRING_STUB
#define PHASE_ORACLE_BOOST_STRING_STUB
#include <algorithm>
#include <cctype>
#include <iostream>
#include <limits>
#include <sstream>
#include <string>
#include <vector>
namespace boost {
namespace algorithm {
inline void to_lower(std::string &s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::tolower(c); }); }
inline void to_upper(std::string &s) { std::transform(s.begin(), s.end(), s.begin(), [](unsigned char c) { return std::toupper(c); }); }
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
inline std::string to_upper_copy(std::string s) { to_upper(s); return s; }
inline void trim(std::string &s) {
  while (!s.empty() && std::isspace((unsigned char)s.back())) s.pop_back();
  size_t i = 0;
  while (i < s.size() && std::isspace((unsigned char)s[i])) ++i;
  s.erase(0, i);
}
inline std::string trim_copy(std::string s) { trim(s); return s; }
struct is_any_of {
  std::string set;
  explicit is_any_of(const std::string &s) : set(s) {}
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
enum token_compress_mode_type { token_compress_on, token_compress_off };
template <class Pred> void split(std::vector<std::string> &out, const std::string &s, Pred p,
                                 token_compress_mode_type mode = token_compress_off) {
  out.clear();
  std::string cur;
  bool lastSep = false;
  for (char c : s) {
    if (p(c)) {
      if (!(mode == token_compress_on && lastSep)) { out.push_back(cur); cur.clear(); }
      lastSep = true;
    } else { cur.push_back(c); lastSep = false; }
  }
  out.push_back(cur);
}
template <class Pred> void trim_if(std::string &s, Pred p) {
  while (!s.empty() && p(s.back())) s.pop_back();
  size_t i = 0;
  while (i < s.size() && p(s[i])) ++i;
  s.erase(0, i);
}
template <class Pred> void trim_left_if(std::string &s, Pred p) { size_t i = 0; while (i < s.size() && p(s[i])) ++i; s.erase(0, i); }
template <class Pred> void trim_right_if(std::string &s, Pred p) { while (!s.empty() && p(s.back())) s.pop_back(); }
inline void erase_all(std::string &s, const std::string &what) {
  if (what.empty()) return;
  for (size_t p = s.find(what); p != std::string::npos; p = s.find(what, p)) s.erase(p, what.size());
}
}  // namespace algorithm
using algorithm::to_lower; using algorithm::to_upper; using algorithm::to_lower_copy; using algorithm::to_upper_copy;
using algorithm::trim; using algorithm::trim_copy; using algorithm::is_any_of; using algorithm::split;
using algorithm::token_compress_on; using algorithm::token_compress_off; using algorithm::trim_if;
using algorithm::trim_left_if; using algorithm::trim_right_if; using algorithm::erase_all;
}  // namespace boost
#endif
