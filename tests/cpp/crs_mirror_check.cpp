// crs_mirror_check.cpp -- replays a scripted sequence of CrsEquation operations on the C++ mirror
// (include/phase/CrsEquation.h) and prints the resulting arrays; the test compares them bit for bit
// with the same sequence run on the reference's own CrsEquation (oracle/_ref).
//   input (stdin): lines "new id n nnz" | "add id r c v" | "set id r c v" | "rhs id r v" | "scale id r v"
//                  | "+= a b" | "-= a b" | "*= a v" | "dump id"
#include <cstdio>
#include <iostream>
#include <map>
#include <sstream>

#include "phase/CrsEquation.h"

int main() {
  std::map<int, CrsEquation> eq;
  std::string line;
  while (std::getline(std::cin, line)) {
    std::istringstream is(line);
    std::string op;
    is >> op;
    if (op == "new") { int id; Size n, nnz; is >> id >> n >> nnz; eq[id] = CrsEquation(n, nnz); }
    else if (op == "add") { int id, r, c; double v; is >> id >> r >> c >> v; eq[id].addCoeff(r, c, v); }
    else if (op == "set") { int id, r, c; double v; is >> id >> r >> c >> v; eq[id].setCoeff(r, c, v); }
    else if (op == "rhs") { int id, r; double v; is >> id >> r >> v; eq[id].addRhs(r, v); }
    else if (op == "scale") { int id, r; double v; is >> id >> r >> v; eq[id].scaleRow(r, v); }
    else if (op == "+=") { int a, b; is >> a >> b; eq[a] += eq[b]; }
    else if (op == "-=") { int a, b; is >> a >> b; eq[a] == eq[b]; }
    else if (op == "*=") { int a; double v; is >> a >> v; eq[a] *= v; }
    else if (op == "dump") {
      int id; is >> id;
      const CrsEquation &e = eq[id];
      printf("rowPtr");
      for (Index v : e.rowPtr()) printf(" %d", v);
      printf("\ncolInd");
      for (Index v : e.colInd()) printf(" %d", v);
      printf("\nvals");
      for (Scalar v : e.vals()) printf(" %.17g", v);
      printf("\nrhs");
      for (Size i = 0; i < e.rank(); ++i) printf(" %.17g", e.b((Index)i));
      printf("\n");
    }
  }
  return 0;
}
