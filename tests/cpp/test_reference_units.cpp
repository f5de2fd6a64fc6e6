// The reference crate's own unit tests for the Neumann / push path, restated against the C++ mirror of its API
// (sublinear-time-solver_b200/cpp/sublinear.hpp). Each block cites the Rust test it follows.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <utility>

#include "../../sublinear-time-solver_b200/cpp/sublinear.hpp"

using namespace sublinear;

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                        \
        }                                                                        \
    } while (0)

int main() {
    {  // matrix/mod.rs:577-587 test_matrix_creation
        auto m = SparseMatrix::from_triplets({{0, 0, 4.0}, {0, 1, 1.0}, {1, 0, 2.0}, {1, 1, 5.0}}, 2, 2);
        CHECK(m.rows() == 2 && m.cols() == 2 && m.nnz() == 4 && m.is_diagonally_dominant());
    }
    {  // matrix/mod.rs:590-600 test_matrix_vector_multiply; sparse.rs:923-933
        auto m = SparseMatrix::from_triplets({{0, 0, 2.0}, {0, 1, 1.0}, {1, 0, 1.0}, {1, 1, 3.0}}, 2, 2);
        std::vector<double> x{1.0, 2.0}, y(2, 0.0);
        m.multiply_vector(x, y);
        CHECK(y[0] == 4.0 && y[1] == 7.0);
    }
    {  // matrix/mod.rs:603-613 test_diagonal_dominance
        CHECK(SparseMatrix::from_triplets({{0, 0, 5.0}, {0, 1, 1.0}, {1, 0, 2.0}, {1, 1, 7.0}}, 2, 2).is_diagonally_dominant());
        CHECK(!SparseMatrix::from_triplets({{0, 0, 1.0}, {0, 1, 3.0}, {1, 0, 2.0}, {1, 1, 2.0}}, 2, 2).is_diagonally_dominant());
    }
    {  // sparse.rs:910-920 test_csr_creation
        auto m = SparseMatrix::from_triplets({{0, 0, 1.0}, {0, 2, 2.0}, {1, 1, 3.0}, {2, 0, 4.0}, {2, 2, 5.0}}, 3, 3);
        CHECK(m.nnz() == 5 && m.get(0, 0) == 1.0 && m.get(0, 2) == 2.0 && m.get(1, 1) == 3.0 && !m.get(0, 1));
    }
    {  // solver/mod.rs:564-581 option presets
        SolverOptions d;
        CHECK(d.tolerance == 1e-6 && d.max_iterations == 1000 && !d.collect_stats);
        CHECK(SolverOptions::fast().tolerance == 1e-3 && SolverOptions::fast().max_iterations == 100);
        CHECK(SolverOptions::high_precision().tolerance == 1e-12 && SolverOptions::high_precision().compute_error_bounds);
    }
    {  // neumann.rs:576-607 test_neumann_solver_simple_system (the intended answer x = [1, 1])
        auto m = SparseMatrix::from_triplets({{0, 0, 4.0}, {0, 1, 1.0}, {1, 0, 1.0}, {1, 1, 3.0}}, 2, 2);
        NeumannSolver s(20, 1e-8);
        auto r = s.solve(m, {5.0, 4.0});
        CHECK(r.converged && std::fabs(r.solution[0] - 1.0) < 0.1 && std::fabs(r.solution[1] - 1.0) < 0.1);
        CHECK(std::fabs(r.solution[0] - 1.0) < 1e-7 && std::fabs(r.solution[1] - 1.0) < 1e-7);
        SolverOptions c;
        c.mode = SolveMode::RefCompat;  // the literal control flow: x_true + D^-1 b, 17 terms
        c.collect_stats = true;
        auto rc = s.solve(m, {5.0, 4.0}, c);
        CHECK(rc.terms_computed == 17 && rc.stats && rc.stats->matvec_count == 21);
        CHECK(std::fabs(rc.solution[0] - 2.25) < 1e-8 && std::fabs(rc.solution[1] - 7.0 / 3.0) < 1e-8);
    }
    {  // neumann.rs:609-631 test_neumann_not_diagonally_dominant
        auto m = SparseMatrix::from_triplets({{0, 0, 1.0}, {0, 1, 3.0}, {1, 0, 2.0}, {1, 1, 1.0}}, 2, 2);
        bool threw = false;
        try {
            NeumannSolver(20, 1e-8).solve(m, {4.0, 3.0});
        } catch (const SolverError &e) {
            threw = e.kind == ErrorKind::MatrixNotDiagonallyDominant && e.is_recoverable();
        }
        CHECK(threw);
    }
    {  // neumann.rs:633-648 test_neumann_state_initialization: diag(2,3) x = [4,6] -> c = [2,2] is already the answer
        auto m = SparseMatrix::from_triplets({{0, 0, 2.0}, {1, 1, 3.0}}, 2, 2);
        auto r = NeumannSolver::default_().solve(m, {4.0, 6.0});
        CHECK(r.solution[0] == 2.0 && r.solution[1] == 2.0);
    }
    {  // from_triplets validation (matrix/mod.rs:166-187)
        bool oob = false, nonfinite = false;
        try { SparseMatrix::from_triplets({{2, 0, 1.0}}, 2, 2); } catch (const SolverError &e) { oob = e.kind == ErrorKind::IndexOutOfBounds; }
        try { SparseMatrix::from_triplets({{0, 0, NAN}}, 2, 2); } catch (const SolverError &e) { nonfinite = e.kind == ErrorKind::InvalidInput; }
        CHECK(oob && nonfinite);
    }
    {  // optimized_solver.rs:380-396 test_optimized_matrix_creation / test_optimized_matrix_vector_multiply
        auto m = OptimizedSparseMatrix::from_triplets({{0, 0, 4.0}, {0, 1, 1.0}, {1, 0, 1.0}, {1, 1, 3.0}}, 2, 2);
        CHECK(m.dimensions() == std::make_pair(size_t(2), size_t(2)) && m.nnz() == 4);
        std::vector<double> x{1.0, 2.0}, y(2, 0.0);
        m.multiply_vector(x, y);
        CHECK(y[0] == 6.0 && y[1] == 7.0);
        CHECK(m.get_performance_stats().first == 1 && m.get_performance_stats().second == 4 * 8 + 2 * 8 + 2 * 8);
        m.reset_stats();
        CHECK(m.get_performance_stats().first == 0);
    }
    {  // optimized_solver.rs:398-435 test_optimized_conjugate_gradient / test_solver_performance_stats
        auto m = OptimizedSparseMatrix::from_triplets({{0, 0, 4.0}, {0, 1, 1.0}, {1, 0, 1.0}, {1, 1, 3.0}}, 2, 2);
        std::vector<double> b{1.0, 2.0};
        OptimizedConjugateGradientSolver solver(OptimizedSolverConfig{});
        auto r = solver.solve(m, b);
        CHECK(r.converged && r.residual_norm < 1e-6 && r.iterations > 0);
        std::vector<double> ax(2, 0.0);
        m.multiply_vector(r.solution, ax);
        CHECK(std::sqrt((ax[0] - b[0]) * (ax[0] - b[0]) + (ax[1] - b[1]) * (ax[1] - b[1])) < 1e-10);
        CHECK(r.performance_stats.matvec_count > 0 && r.performance_stats.dot_product_count > 0 &&
              r.performance_stats.total_flops > 0);
        CHECK(solver.get_last_iteration_count() == r.performance_stats.matvec_count && r.data().size() == 2);
        bool threw = false;  // "Right-hand side vector length must match matrix size" (:191-193)
        try { solver.solve(m, {1.0, 2.0, 3.0}); } catch (const SolverError &e) { threw = e.kind == ErrorKind::DimensionMismatch; }
        CHECK(threw);
    }
    std::printf("reference unit tests: all passed\n");
    return 0;
}
