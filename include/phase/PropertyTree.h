// PropertyTree.h -- minimal stand-in for boost::property_tree::ptree (Boost is
// not a dependency of this build) with the calls the hot path's callers make:
// get<T>(path), get<T>(path, default), get_child, get_child_optional, put, and a
// reader for the INFO format of the case files (S/Input.cpp:8-15).
#ifndef PHASE_B200_PROPERTY_TREE_H
#define PHASE_B200_PROPERTY_TREE_H
#include <fstream>
#include <memory>
#include <sstream>
#include <string>
#include <utility>
#include <vector>

#include "Exception.h"

namespace phase {
class PropertyTree {
public:
  typedef std::vector<std::pair<std::string, PropertyTree>> Children;
  PropertyTree() {}
  explicit PropertyTree(const std::string &data) : data_(data) {}

  const std::string &data() const { return data_; }
  Children::const_iterator begin() const { return children_.begin(); }
  Children::const_iterator end() const { return children_.end(); }
  bool empty() const { return children_.empty(); }

  const PropertyTree *find(const std::string &path) const {
    const PropertyTree *t = this;
    size_t pos = 0;
    while (pos <= path.size()) {
      const size_t dot = path.find('.', pos);
      const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
      const PropertyTree *next = nullptr;
      for (const auto &c : t->children_)
        if (c.first == key) { next = &c.second; break; }
      if (!next) return nullptr;
      t = next;
      if (dot == std::string::npos) break;
      pos = dot + 1;
    }
    return t;
  }
  const PropertyTree &get_child(const std::string &path) const {
    const PropertyTree *t = find(path);
    if (!t) throw Exception("ptree", "get_child", "No such node (" + path + ")");
    return *t;
  }
  const PropertyTree *get_child_optional(const std::string &path) const { return find(path); }

  template <class T> T get(const std::string &path) const {
    const PropertyTree *t = find(path);
    if (!t) throw Exception("ptree", "get", "No such node (" + path + ")");
    return convert<T>(t->data_);
  }
  template <class T> T get(const std::string &path, const T &def) const {
    const PropertyTree *t = find(path);
    return t ? convert<T>(t->data_) : def;
  }
  std::string get(const std::string &path, const char *def) const { return get<std::string>(path, std::string(def)); }

  PropertyTree &put(const std::string &path, const std::string &value) {
    PropertyTree *t = this;
    size_t pos = 0;
    for (;;) {
      const size_t dot = path.find('.', pos);
      const std::string key = path.substr(pos, dot == std::string::npos ? std::string::npos : dot - pos);
      PropertyTree *next = nullptr;
      for (auto &c : t->children_)
        if (c.first == key) { next = &c.second; break; }
      if (!next) {
        t->children_.push_back(std::make_pair(key, PropertyTree()));
        next = &t->children_.back().second;
      }
      t = next;
      if (dot == std::string::npos) break;
      pos = dot + 1;
    }
    t->data_ = value;
    return *t;
  }
  template <class T> PropertyTree &put(const std::string &path, const T &value) {
    std::ostringstream os;
    os.precision(17);
    os << value;
    return put(path, os.str());
  }

  // INFO format: `key value`, `key { ... }`, `; comment`, quoted strings
  static PropertyTree parseInfo(std::istream &in) {
    PropertyTree root;
    std::vector<PropertyTree *> stack(1, &root);
    std::string line, lastKey;
    while (std::getline(in, line)) {
      std::vector<std::string> tok = tokenize(line);
      for (size_t i = 0; i < tok.size(); ++i) {
        if (tok[i] == "{") {
          PropertyTree *cur = stack.back();
          if (lastKey.empty() || cur->children_.empty()) throw Exception("ptree", "parseInfo", "unexpected {");
          stack.push_back(&cur->children_.back().second);
          lastKey.clear();
        } else if (tok[i] == "}") {
          if (stack.size() == 1) throw Exception("ptree", "parseInfo", "unmatched }");
          stack.pop_back();
          lastKey.clear();
        } else {
          const std::string key = tok[i];
          std::string val;
          if (i + 1 < tok.size() && tok[i + 1] != "{" && tok[i + 1] != "}") val = tok[++i];
          stack.back()->children_.push_back(std::make_pair(key, PropertyTree(val)));
          lastKey = key;
        }
      }
    }
    if (stack.size() != 1) throw Exception("ptree", "parseInfo", "unmatched {");
    return root;
  }
  static PropertyTree readInfo(const std::string &filename) {
    std::ifstream f(filename);
    if (!f) throw Exception("ptree", "readInfo", "cannot open " + filename);
    return parseInfo(f);
  }

private:
  static std::vector<std::string> tokenize(const std::string &line) {
    std::vector<std::string> out;
    size_t i = 0;
    while (i < line.size()) {
      const char c = line[i];
      if (c == ';') break;
      if (c == ' ' || c == '\t' || c == '\r') { ++i; continue; }
      if (c == '{' || c == '}') { out.push_back(std::string(1, c)); ++i; continue; }
      if (c == '"') {
        const size_t e = line.find('"', i + 1);
        out.push_back(line.substr(i + 1, e == std::string::npos ? std::string::npos : e - i - 1));
        i = e == std::string::npos ? line.size() : e + 1;
        continue;
      }
      size_t e = i;
      while (e < line.size() && line[e] != ' ' && line[e] != '\t' && line[e] != '\r' && line[e] != '{' &&
             line[e] != '}' && line[e] != ';')
        ++e;
      out.push_back(line.substr(i, e - i));
      i = e;
    }
    return out;
  }
  template <class T> static T convert(const std::string &s) {
    std::istringstream is(s);
    T v;
    is >> v;
    if (is.fail()) throw Exception("ptree", "get", "conversion of \"" + s + "\" failed");
    return v;
  }
  std::string data_;
  Children children_;
};
template <> inline std::string PropertyTree::convert<std::string>(const std::string &s) { return s; }
}  // namespace phase

// the reference spells the type boost::property_tree::ptree
namespace boost { namespace property_tree { typedef ::phase::PropertyTree ptree; } }
#endif
