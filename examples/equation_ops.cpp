// equation_ops.cpp -- the parts of the FiniteVolumeEquation<T> interface beyond the fractional-step statements,
// through the mirrored API (include/phase): the cell-group operators fv::ddt(field, dt, cells) and
// src::div(field, cells), src::laplacian, and the per-entry Vector2D / Tensor2D coefficients of a vector equation
// (UE/VectorFiniteVolumeEquation.cpp:38-76) solved through Seam 1.
//   usage: equation_ops <out.bin>
// Writes, as raw doubles: [n, nnz1, rowPtr1.., colInd1.., vals1.., rhs1.., ux.., uy..] for the parity test.
#include <cmath>
#include <cstdio>
#include <valarray>

#include "phase/FiniteVolumeEquation.h"

int main(int argc, char *argv[]) {
  try {
    auto comm = std::make_shared<const Communicator>(0);
    std::shared_ptr<const FiniteVolumeGrid2D> grid = std::make_shared<StructuredRectilinearGrid>(comm, 12, 9, 1.2, 0.9);
    VectorFiniteVolumeField u(grid, "u", Vector2D(0., 0.));
    ScalarFiniteVolumeField p(grid, "p", 0.), phi(grid, "phi", 0.);
    for (const Cell &c : grid->cells()) {
      const Point2D &x = c.centroid();
      u(c) = Vector2D(std::sin(3. * x.x) + x.y, std::cos(2. * x.y) * x.x);
      p(c) = x.x * x.x + 0.5 * x.y;
      phi(c) = std::cos(x.x + 2. * x.y);
    }
    u.interpolateFaces();
    p.setBoundaryFaces();
    phi.savePreviousTimeStep(0.01, 1);

    CellGroup grp("everyThird");
    for (const Cell &c : grid->cells())
      if (c.id() % 3 == 0) grp.add(c);

    // (1) operator expression with the cell-group overloads and src::laplacian
    FiniteVolumeEquation<Scalar> phiEqn(phi, "phiEqn");
    phiEqn = (fv::ddt(phi, 0.01, grp) == src::div(u, grp) - src::laplacian(0.7, p));
    std::vector<Index> rowPtr, colInd;
    std::vector<Scalar> vals, rhs;
    phiEqn.exportReferenceLayout(0, rowPtr, colInd, vals, rhs);

    // (2) per-entry vector equation: a_P = tensor with off-diagonal coupling, neighbours as Vector2D
    boost::property_tree::ptree params;
    params.put("lib", "b200");
    params.put("solver", "BICGSTAB");
    params.put("preconditioner", "jacobi");
    params.put("maxIters", 2000);
    params.put("tolerance", 1e-13);
    std::shared_ptr<SparseMatrixSolver> solver = SparseMatrixSolverFactory().create("b200", *comm);
    solver->setup(params);
    FiniteVolumeEquation<Vector2D> uEqn(u, "uEqn", 10);
    uEqn.setSparseSolver(solver);
    for (const Cell &c : grid->cells()) {
      const Point2D &x = c.centroid();
      Tensor2D aP(6. + x.x, 0.25 * x.y, c.id() % 2 ? -0.5 : 0., 7. - x.y);  // yx == 0 on even cells: not stored
      uEqn.add(c, c, aP);
      std::vector<Ref<const Cell>> nbs;
      std::vector<Scalar> coeffs;
      for (const Cell &nb : grid->cells())
        if (std::abs((long)nb.id() - (long)c.id()) == 12) { nbs.push_back(std::cref(nb)); coeffs.push_back(-0.75); }
      uEqn.add(c, nbs, std::valarray<Scalar>(coeffs.data(), coeffs.size()));
      if (c.id() % 12) uEqn.add(c, grid->cells()[c.id() - 1], Vector2D(-1., -1.25));
      uEqn.addSource(c, Vector2D(-std::sin(x.x), 1. + x.y));
    }
    const Cell &probe = grid->cells()[17];
    const Vector2D d = uEqn.get(probe, probe), w = uEqn.get(probe, grid->cells()[16]);
    printf("get(17,17) = %.17g %.17g   get(17,16) = %.17g %.17g\n", d.x, d.y, w.x, w.y);
    uEqn.solve();
    printf("uEqn iterations %d\n", solver->nIters());

    bool refused = false;
    try { phiEqn.add(probe, probe, Tensor2D(1., 0., 0., 1.)); } catch (const Exception &) { refused = true; }
    if (!refused) return 2;

    if (argc > 1) {
      FILE *f = fopen(argv[1], "wb");
      std::vector<double> out;
      const Size n = grid->nCells();
      out.push_back((double)n);
      out.push_back((double)colInd.size());
      for (Index v : rowPtr) out.push_back((double)v);
      for (Index v : colInd) out.push_back((double)v);
      out.insert(out.end(), vals.begin(), vals.end());
      out.insert(out.end(), rhs.begin(), rhs.end());
      for (const Cell &c : grid->cells()) out.push_back(u(c).x);
      for (const Cell &c : grid->cells()) out.push_back(u(c).y);
      fwrite(out.data(), sizeof(double), out.size(), f);
      fclose(f);
    }
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
