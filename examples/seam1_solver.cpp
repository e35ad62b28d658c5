// seam1_solver.cpp -- Seam 1 exactly as the reference drives it: a host CrsEquation
// filled by addCoeff/addRhs, a backend made by SparseMatrixSolverFactory from the
// `lib` string, setup(ptree), then CrsEquation::solve() = setRank + set(rowPtr,
// colInd, vals) + setRhs(-rhs_) + solve, and x(i) read back per element.
//   usage: seam1_solver <n>     (n x n 5-point Poisson, Dirichlet, ELL-5 padded rows)
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "phase/CrsEquation.h"
#include "phase/SparseMatrixSolverFactory.h"

int main(int argc, char *argv[]) {
  const int n = argc > 1 ? atoi(argv[1]) : 32;
  try {
    Communicator comm(0);
    boost::property_tree::ptree params;
    params.put("lib", "b200");
    params.put("solver", "BICGSTAB");
    params.put("preconditioner", "jacobi");
    params.put("maxIters", 5000);
    params.put("tolerance", 1e-10);
    std::shared_ptr<SparseMatrixSolver> solver = SparseMatrixSolverFactory().create(params.get<std::string>("lib"), comm);
    solver->setup(params);
    CrsEquation eqn((Size)n * n, 5);
    eqn.setSparseSolver(solver);
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        const Index r = j * n + i;
        if (i > 0) eqn.addCoeff(r, r - 1, 1.);
        eqn.addCoeff(r, r, -4.);
        if (i + 1 < n) eqn.addCoeff(r, r + 1, 1.);
        if (j > 0) eqn.addCoeff(r, r - n, 1.);
        if (j + 1 < n) eqn.addCoeff(r, r + n, 1.);
        eqn.addRhs(r, std::sin(0.1 * r));  // A x + rhs = 0
      }
    eqn.solve();
    // residual of A x = -rhs with the host copy
    double rr = 0., bb = 0.;
    for (Index r = 0; r < n * n; ++r) {
      double ax = 0.;
      for (Index j = eqn.rowPtr()[r]; j < eqn.rowPtr()[r + 1]; ++j)
        if (eqn.colInd()[j] >= 0) ax += eqn.vals()[j] * eqn.x(eqn.colInd()[j]);
      rr += (ax + eqn.b(r)) * (ax + eqn.b(r));
      bb += eqn.b(r) * eqn.b(r);
    }
    solver->printStatus("CrsEquation seam1:");
    printf("relres %.3e iters %d\n", std::sqrt(rr / bb), solver->nIters());
    try {
      SparseMatrixSolverFactory().create("eigen", comm);
      return 2;
    } catch (const Exception &e) {
      printf("refused: %s\n", e.what());
    }
    return std::sqrt(rr / bb) <= 1.01e-10 ? 0 : 1;
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
