// lid_driven_cavity.cpp -- the time loop of S/RunControl.cpp:21-52 around the
// FractionalStep module, entirely through the mirrored Phase API (include/phase).
//   usage: lid_driven_cavity <caseDir> <nSteps> <out.bin>
// Writes u.x, u.y, p (cells) as raw doubles for the parity test.
#include <cstdio>
#include <cstdlib>

#include "phase/FractionalStep.h"

int main(int argc, char *argv[]) {
  const std::string caseDir = argc > 1 ? argv[1] : "case";
  const int nSteps = argc > 2 ? atoi(argv[2]) : 10;
  try {
    Input input(caseDir);
    input.parseInputFile();
    auto comm = std::make_shared<const Communicator>(0);
    std::shared_ptr<const FiniteVolumeGrid2D> grid = std::make_shared<StructuredRectilinearGrid>(comm, input);
    FractionalStep solver(input, grid);
    solver.initialize();
    const Scalar dt = input.caseInput().get<Scalar>("Solver.timeStep");
    for (int i = 0; i < nSteps; ++i) solver.solve(dt);
    if (argc > 3) {
      FILE *f = fopen(argv[3], "wb");
      const Size n = grid->nCells();
      std::vector<double> out(3 * n);
      for (const Cell &c : grid->cells()) {
        out[c.id()] = solver.u()(c).x;
        out[n + c.id()] = solver.u()(c).y;
        out[2 * n + c.id()] = solver.p()(c);
      }
      fwrite(out.data(), sizeof(double), out.size(), f);
      fclose(f);
    }
    printf("done: %d steps, %zu cells\n", nSteps, (size_t)grid->nCells());
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
