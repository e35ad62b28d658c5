// phase_piso.cpp -- config 1: Examples/LidDrivenCavity (100x100, Re = 10) with the PISO module and
// the legacy case keys, through the mirrored C++ API.
//   usage: phase_piso <caseDir> [maxIterations] [out.bin]
#include <cstdio>
#include <cstdlib>

#include "phase/Piso.h"

int main(int argc, char *argv[]) {
  try {
    Input input(argc > 1 ? argv[1] : "case");
    input.parseInputFile();
    auto comm = std::make_shared<const Communicator>(0);
    std::shared_ptr<const FiniteVolumeGrid2D> grid = std::make_shared<StructuredRectilinearGrid>(comm, input);
    Piso solver(input, grid);
    solver.initialize();
    const Scalar dt = input.caseInput().get<Scalar>("Solver.timeStep");
    const int n = argc > 2 ? atoi(argv[2]) : input.caseInput().get<int>("Solver.maxIterations", 100);
    Scalar m = 0.;
    for (int i = 0; i < n; ++i) m = solver.solve(dt);
    if (argc > 3) {
      const std::vector<double> u = solver.field("u"), p = solver.field("p");
      FILE *f = fopen(argv[3], "wb");
      fwrite(u.data(), sizeof(double), u.size(), f);
      fwrite(p.data(), sizeof(double), p.size(), f);
      fclose(f);
    }
    printf("done: %d iterations, final mass imbalance %.3e\n", n, m);
    return m < 1e-6 ? 0 : 1;
  } catch (const std::exception &e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
