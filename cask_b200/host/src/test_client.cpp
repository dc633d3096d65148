// The reference's gtest suites for the path, as one self-contained binary (gtest is an empty submodule
// in the reference tree): test/LinearSolvers.cpp:14-52, test/CgTest.cpp:34-51, test/ClientTestSpmv.cpp:10-26,
// test/ClientTestCg.cpp:8-21 — the two client tests end in ASSERT_TRUE(false) upstream ("make test fail
// until we implement it"); here they assert real results.
//   usage: test_client <dir with tiny.mtx tiny_b.mtx tiny_sol.mtx tinysym.mtx tinysym_b.mtx tinysym_sol.mtx>
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "Benchmark.hpp"
#include "Cask.hpp"
#include "IO.hpp"
#include "SparseLinearSolvers.hpp"

namespace {
int g_failures = 0;
std::string g_dir;

#define CHECK(cond)                                                                       \
  do {                                                                                    \
    if (!(cond)) { std::cerr << __FILE__ << ":" << __LINE__ << ": " #cond << std::endl; g_failures++; } \
  } while (0)

// ASSERT_DOUBLE_EQ: within 4 units in the last place
bool double_eq(double a, double b) {
  if (a == b) return true;
  int64_t ia, ib;
  std::memcpy(&ia, &a, 8);
  std::memcpy(&ib, &b, 8);
  if ((ia < 0) != (ib < 0)) return false;
  return std::llabs(ia - ib) <= 4;
}

using namespace cask;
using namespace cask::sparse_linear_solvers;

void CGWithIdentityPC() {  // test/LinearSolvers.cpp:14-32
  Vector rhs = io::readVector(g_dir + "/tiny_b.mtx");
  SymCsrMatrix a = io::readSymMatrix(g_dir + "/tiny.mtx");
  int iterations = 0;
  Vector sol(a.n);
  bool ok = pcg<double, IdentityPreconditioner>(a.matrix, &rhs[0], &sol[0], iterations);
  Vector exp_sol{1, 2, 3, 4};
  sol.print("got x = ");
  std::cout << "Iterations = " << iterations << std::endl;
  CHECK(ok);
  for (int i = 0; i < sol.size(); i++) CHECK(double_eq(sol[i], exp_sol[i]));
}

void CGSymWithIdentityPC() {  // test/LinearSolvers.cpp:34-52
  Vector rhs = io::readVector(g_dir + "/tinysym_b.mtx");
  SymCsrMatrix a = io::readSymMatrix(g_dir + "/tinysym.mtx");
  int iterations = 0;
  Vector sol(a.n);
  bool ok = pcg<>(a.matrix, &rhs[0], &sol[0], iterations);
  Vector exp_sol{-2, 2, 3, 3};
  sol.print("got x =");
  std::cout << "Iterations = " << iterations << std::endl;
  CHECK(ok);
  for (int i = 0; i < sol.size(); i++) CHECK(double_eq(sol[i], exp_sol[i]));
}

void CGSymWithILUPC() {  // test/LinearSolvers.cpp:54-77
  Vector rhs = io::readVector(g_dir + "/tinysym_b.mtx");
  SymCsrMatrix a = io::readSymMatrix(g_dir + "/tinysym.mtx");
  int iterations = 0;
  Vector sol(a.n);
  pcg<double, ILUPreconditioner>(a.matrix, &rhs[0], &sol[0], iterations);  // the reference ignores the return value too
  Vector exp_sol{-1.9982580059252246, 2.0000862488691915, 3.0001293733037859, 2.9987581910958183};
  sol.print("got x =");
  std::cout << "Iterations = " << iterations << std::endl;
  // the loop stalls (non-symmetric M) and x is where it stands after 2000 iterations; the GPU's dot products sum in
  // another order than MKL's, so the asserted doubles are met to ~1e-12 rather than to 4 ulp
  for (int i = 0; i < sol.size(); i++) CHECK(std::fabs(sol[i] - exp_sol[i]) <= 1e-9 * std::fabs(exp_sol[i]));
}

void ILUCompute2() {  // test/LinearSolvers.cpp:79-99
  CsrMatrix a{2, 1, 1, 1, 1, 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 1};
  ILUPreconditioner ilupc{a};
  DokMatrix exp{2, 1, 1, 1, 0.5, 0.5, 0, 0, 0.5, 0, 0.5, 0, 0.5, 0, 0, 0.5};
  CHECK(ilupc.pc.n == exp.n);
  CHECK(ilupc.pc.nnzs == exp.nnzs);
  CHECK(ilupc.pc == exp);
}

void ILUCompute() {  // test/LinearSolvers.cpp:101-123
  SymCsrMatrix a = io::readSymMatrix(g_dir + "/tinysym.mtx");
  CsrMatrix explicitA(a.matrix.toDok().explicitSymmetric());
  ILUPreconditioner explicitPc{explicitA};
  CsrMatrix csrPc{explicitPc.pc};
  CHECK((csrPc.row_ptr == std::vector<int>{0, 2, 3, 4, 6}));
  CHECK((csrPc.col_ind == std::vector<int>{0, 3, 1, 2, 0, 3}));
  CHECK((csrPc.values == std::vector<double>{1, 1, 1, 1, 1, 1}));
}

void ILUComputeAndApply() {  // test/LinearSolvers.cpp:125-146
  ILUPreconditioner ilupc{CsrMatrix{DokMatrix{2, 1, 1, 1, 1, 1, 0, 0, 1, 0, 1, 0, 1, 0, 0, 1}}};
  std::vector<double> v{1, 2, 3, 4};
  auto res = ilupc.apply(v);
  std::vector<double> expPcApply{-16.25, 7, 11, 15};
  for (size_t i = 0; i < res.size(); i++) CHECK(double_eq(res[i], expPcApply[i]));
}

void PreconditionersBeyondTheReference() {  // Jacobi converges to the reference's solution; the unit-lower ILU follows the oracle
  Vector rhs = io::readVector(g_dir + "/tinysym_b.mtx");
  SymCsrMatrix a = io::readSymMatrix(g_dir + "/tinysym.mtx");
  Vector exp_sol{-2, 2, 3, 3};
  {
    int iterations = 0;
    Vector sol(a.n);
    const bool ok = pcg<double, JacobiPreconditioner>(a.matrix, &rhs[0], &sol[0], iterations);
    CHECK(ok);
    for (int i = 0; i < sol.size(); i++) CHECK(std::fabs(sol[i] - exp_sol[i]) < 1e-6);
  }
  {
    // pcg<> builds its preconditioner from the arrays it is HANDED, the stored lower triangle (:171), so the factors are
    // those of a triangular matrix and M is not symmetric whichever way the lower solve is applied: like the reference's
    // own CGSymWithILUPC the loop runs out of iterations (oracle_pcg_precond, PRECON_ILU_UNIT: not converged, 1999,
    // x = {-2.0982, 1.9875, 2.9812, 3.0366})
    int iterations = 0;
    Vector sol(a.n);
    const bool ok = pcg<double, IluUnitPreconditioner>(a.matrix, &rhs[0], &sol[0], iterations);
    CHECK(!ok);
    CHECK(iterations == 1999);
    const double oracle_x[4] = {-2.09821451, 1.9874614, 2.9811921, 3.03656865};
    for (int i = 0; i < sol.size(); i++) CHECK(std::fabs(sol[i] - oracle_x[i]) < 1e-3);
  }
}

void GpuReadersMatchTheHostReaders() {  // io::gpu::readMatrix / readSymMatrix against io::readMatrix / readSymMatrix
  for (const char* name : {"tiny", "tinysym"}) {
    const std::string p = g_dir + "/" + name + ".mtx";
    CHECK(io::gpu::readMatrix(p) == io::readMatrix(p));
    CHECK(io::gpu::readSymMatrix(p).matrix == io::readSymMatrix(p).matrix);
  }
}

void CgTestRun(const std::string& name) {  // test/CgTest.cpp:10-43 (identity preconditioner leg)
  SymCsrMatrix a = io::readSymMatrix(g_dir + "/" + name + ".mtx");
  Vector rhs = io::readVector(g_dir + "/" + name + "_b.mtx");
  Vector exp = io::readVector(g_dir + "/" + name + "_sol.mtx");
  int iterations = 0;
  Vector sol(a.n);
  utils::Timer t;
  pcg<double, IdentityPreconditioner>(a.matrix, &rhs.data[0], &sol.data[0], iterations, false, &t);
  std::ofstream log(g_dir + "/sol.upc." + name + ".log");
  benchmark::printSummary(t.get("cg:setup").count(), iterations, t.get("cg:solve").count(), sol.distance(rhs),
                          exp.distance(sol), 0, log);
  sol.writeToFile(g_dir + "/sol.upc." + name + ".mtx");
  CHECK(exp.distance(sol) < 1e-9);
}

void ClientTestSpmv() {  // test/ClientTestSpmv.cpp:10-26
  CaskContext cc;
  CsrMatrix a = io::readMatrix(g_dir + "/tinysym.mtx");
  Vector rhs = io::readVector(g_dir + "/tinysym_b.mtx");
  Vector v(rhs);
  auto spmv = cc.getSpmv(a);
  spmv.preprocess(a);
  Vector got = spmv.spmv(v);
  got.print("Spmv res = ");
  Vector exp = a.dot(v);  // the reference's own CsrMatrix::dot as the expected value
  CHECK(got == exp);
  // A * (solution of A x = b) == b
  Vector sol = io::readVector(g_dir + "/tinysym_sol.mtx");
  CHECK(spmv.spmv(sol) == rhs);
  // north_star aliases
  cask::spmv::SimpleSpmvArchitecture& alias = spmv;
  CHECK(alias.dfespmv(v) == exp);
  CHECK(spmv.get_name() == "Simple");
  CHECK(spmv.getEstimatedClockCycles() > 0);
}

void ClientCg() {  // test/ClientTestCg.cpp:8-21
  CaskContext cc;
  SymCsrMatrix a = io::readSymMatrix(g_dir + "/tiny.mtx");
  Vector rhs = io::readVector(g_dir + "/tiny_b.mtx");
  Vector v(rhs);
  solvers::Cg cg = cc.getCg(a);
  cg.preprocess(a);
  Vector res = cg.solve(v);
  Vector exp = io::readVector(g_dir + "/tiny_sol.mtx");
  CHECK(cg.converged);
  for (int i = 0; i < res.size(); i++) CHECK(double_eq(res[i], exp[i]));
}

void ErrorsMirrorReference() {  // Spmv.cpp:195-232
  CsrMatrix a = io::readMatrix(g_dir + "/tinysym.mtx");
  Vector x(a.m);
  cask::spmv::Spmv small(64, 4, 1, /*maxRows=*/2, 1);
  small.preprocess(a);
  try { small.spmv(x); CHECK(false); } catch (std::invalid_argument& e) {
    CHECK(std::string(e.what()) == "Matrix is too large! Maximum supported rows: 2 actual rows: 4");
  }
  cask::spmv::Spmv odd(64, 4, 3, 100, 2);
  odd.preprocess(a);
  try { odd.spmv(x); CHECK(false); } catch (std::runtime_error& e) {
    CHECK(std::string(e.what()) == "numPipes should be a multiple of numControllers");
  }
  cask::spmv::SkipEmptyRowsSpmv skip(2, 2, 2, 100, 1);
  CHECK(skip.get_name() == "SkipEmpty");
  auto part = skip.do_blocking(a, 2, 2);
  CHECK(part.nBlocks == 2 && part.n == 4 && (int)part.m_colptr.size() == part.m_colptr_unpaddedLength);
}

}  // namespace

int main(int argc, char** argv) {
  if (argc != 2 && argc != 3) { std::cerr << "usage: test_client <systems dir> [base|extra|all]" << std::endl; return 2; }
  g_dir = argv[1];
  // "base": the suites that have been green on a B200 since round 1 (default, what tests/test_gpu_harness.py runs);
  // "extra": the ILU / Jacobi / GPU-ingest suites added afterwards (tests/test_gpu_w_precond.py runs them)
  const std::string which = argc == 3 ? argv[2] : "base";
  struct { const char* name; void (*fn)(); int extra; } tests[] = {
      {"TestLinearSolvers.CGWithIdentityPC", CGWithIdentityPC, 0}, {"TestLinearSolvers.CGSymWithIdentityPC", CGSymWithIdentityPC, 0},
      {"CgTest.SolveTiny", [] { CgTestRun("tiny"); }, 0}, {"CgTest.SolveTinySym", [] { CgTestRun("tinysym"); }, 0},
      {"ClientTestSpmv.TinySymSpmv", ClientTestSpmv, 0}, {"ClientCg.SimpleSystem", ClientCg, 0},
      {"Spmv.ErrorsMirrorReference", ErrorsMirrorReference, 0},
      {"TestLinearSolvers.CGSymWithILUPC", CGSymWithILUPC, 1}, {"TestLinearSolvers.ILUCompute2", ILUCompute2, 1},
      {"TestLinearSolvers.ILUCompute", ILUCompute, 1}, {"TestLinearSolvers.ILUComputeAndApply", ILUComputeAndApply, 1},
      {"Precon.JacobiAndUnitIlu", PreconditionersBeyondTheReference, 1}, {"Io.GpuReaders", GpuReadersMatchTheHostReaders, 1}};
  for (auto& t : tests) {
    if ((which == "base" && t.extra) || (which == "extra" && !t.extra)) continue;
    const int before = g_failures;
    std::cout << "[ RUN      ] " << t.name << std::endl;
    try { t.fn(); } catch (std::exception& e) { std::cerr << "exception: " << e.what() << std::endl; g_failures++; }
    std::cout << (g_failures == before ? "[       OK ] " : "[  FAILED  ] ") << t.name << std::endl;
  }
  std::cout << (g_failures ? "FAILED" : "PASSED") << " (" << g_failures << " failures)" << std::endl;
  return g_failures ? 1 : 0;
}
