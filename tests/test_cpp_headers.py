"""The C++ mirror of the Phase API (include/phase) compiles stand-alone with g++ (no GPU needed)
and links against libphase_b200.so."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_headers_compile_standalone(tmp_path):
    hdrs = sorted(f for f in os.listdir(os.path.join(ROOT, "include", "phase")) if f.endswith(".h"))
    assert "FiniteVolumeEquation.h" in hdrs and "B200SparseMatrixSolver.h" in hdrs
    for h in hdrs:
        src = tmp_path / ("t_" + h.replace(".h", ".cpp"))
        src.write_text('#include "phase/%s"\nint main() { return 0; }\n' % h)
        r = subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), "-fsyntax-only", str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, h + "\n" + r.stderr


def test_examples_link():
    from examples import build as eb
    for exe in eb.build():
        assert os.path.exists(exe)


def test_info_parser_reads_reference_style_case(tmp_path):
    """PropertyTree parses the INFO format of the case files (comments, nesting, quoted strings, vectors)."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <cstdio>
#include "phase/Input.h"
#include "phase/Vector2D.h"
int main(int argc, char** argv) {
  Input in(argv[1]); in.parseInputFile();
  const auto& c = in.caseInput();
  if (c.get<std::string>("Solver.type") != "fractional step") return 1;
  if (c.get<double>("Solver.timeStep") != 5e-3) return 2;
  if (c.get<std::string>("LinearAlgebra.pEqn.lib") != "b200") return 3;
  if (c.get<int>("Grid.nCellsX") != 40) return 4;
  if (c.get<double>("Properties.missing", 7.5) != 7.5) return 5;
  Vector2D v(in.boundaryInput().get<std::string>("Boundaries.u.y+.value"));
  if (v.x != 1. || v.y != 0.) return 6;
  if (in.boundaryInput().get<std::string>("Boundaries.p.*.type") != "normal_gradient") return 7;
  try { c.get<int>("Nope.nope"); return 8; } catch (const Exception&) {}
  return 0;
}''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++14", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    r = subprocess.run([str(exe), os.path.join(ROOT, "examples", "LidDrivenCavity", "case")])
    assert r.returncode == 0, r.returncode
