// Stand-in for Boost.PropertyTree's INFO reader (TEST INFRASTRUCTURE, see ptree.hpp): `key value`, `key { ... }`,
// `; comment`, quoted strings -- the subset the reference's case files use.
#ifndef PHASE_ORACLE_INFO_PARSER_STUB
#define PHASE_ORACLE_INFO_PARSER_STUB
#include <fstream>
#include <boost/property_tree/ptree.hpp>
namespace boost { namespace property_tree {
namespace info_parser {
inline std::vector<std::string> tokens(const std::string &line) {
  std::vector<std::string> out;
  std::size_t i = 0;
  while (i < line.size()) {
    const char c = line[i];
    if (c == ';') break;
    if (c == ' ' || c == '\t' || c == '\r') { ++i; continue; }
    if (c == '{' || c == '}') { out.push_back(std::string(1, c)); ++i; continue; }
    if (c == '"') {
      const std::size_t e = line.find('"', i + 1);
      out.push_back(line.substr(i + 1, e == std::string::npos ? std::string::npos : e - i - 1));
      i = e == std::string::npos ? line.size() : e + 1;
      continue;
    }
    std::size_t e = i;
    while (e < line.size() && std::string(" \t\r{};").find(line[e]) == std::string::npos) ++e;
    out.push_back(line.substr(i, e - i));
    i = e;
  }
  return out;
}
inline void read_info(std::istream &in, ptree &root) {
  root = ptree();
  std::vector<ptree *> stack(1, &root);
  std::string line;
  bool haveKey = false;
  while (std::getline(in, line)) {
    const std::vector<std::string> tok = tokens(line);
    for (std::size_t i = 0; i < tok.size(); ++i) {
      if (tok[i] == "{") {
        if (!haveKey) throw ptree_error("info: unexpected {");
        stack.push_back(&stack.back()->children().back().second);
        haveKey = false;
      } else if (tok[i] == "}") {
        if (stack.size() == 1) throw ptree_error("info: unmatched }");
        stack.pop_back();
        haveKey = false;
      } else {
        const std::string key = tok[i];
        std::string val;
        if (i + 1 < tok.size() && tok[i + 1] != "{" && tok[i + 1] != "}") val = tok[++i];
        stack.back()->children().push_back(std::make_pair(key, ptree(val)));
        haveKey = true;
      }
    }
  }
  if (stack.size() != 1) throw ptree_error("info: unmatched {");
}
inline void read_info(const std::string &filename, ptree &root) {
  std::ifstream f(filename.c_str());
  if (!f) throw ptree_error("cannot open " + filename);
  read_info(f, root);
}
}  // namespace info_parser
using info_parser::read_info;
}}  // namespace boost::property_tree
#endif
