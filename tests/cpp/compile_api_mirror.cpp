// Compile-and-link check (no GPU needed, nothing is executed against a device): every member of the C++ mirror of the
// reference API is instantiated so that a signature drifting away from include/sublinear_b200.h breaks the build.
#include "../../sublinear-time-solver_b200/cpp/sublinear.hpp"

using namespace sublinear;

int use_everything(bool run) {
    if (!run) return 0;  // never true in the test: the point is that this translation unit compiles and links
    auto m = SparseMatrix::from_triplets({{0, 0, 4.0}, {0, 1, 1.0}, {1, 0, 1.0}, {1, 1, 3.0}}, 2, 2);
    NeumannSolver solver(20, 1e-8);
    NeumannState st = solver.initialize(m, {5.0, 4.0});
    while (solver.step(st) == StepResult::Continue && !solver.is_converged(st)) {}
    solver.update_rhs(st, {{0, 0.5}});
    st.reset();
    auto x = solver.extract_solution(st);
    (void)st.residual_norm(); (void)st.matvec_count(); (void)st.error_bounds(); (void)st.memory_usage();
    (void)st.terms_computed(); (void)st.series_converged();
    PushGraph g = PushGraph::from_edges(3, {{0, 1, 0.5}, {1, 2, 1.0}, {2, 0, 0.3}});
    (void)g.num_edges(); (void)g.out_degree(0); (void)g.in_degree(0);
    ForwardPushSolver f(PushGraph::from_matrix({0, 1, 2}, {1, 0}, {1.0, 1.0}), PushConfig{});
    auto fr = f.solve_single_source(0);
    (void)f.solve_multi_source({0, 1}); (void)f.query_single_entry(0, 1); (void)f.extrapolated_solution(fr);
    BackwardPushSolver b(std::move(g));
    auto br = b.solve_single_target(0);
    (void)b.solve_multi_target({0, 1}); (void)b.query_transition_probability(0, 1);
    (void)f.solve_with_target(0, 1, 1e-3); (void)b.solve_with_source(0, 1, 1e-3);
    (void)b.extrapolated_solution(br); (void)b.reachability_probabilities(1);
    (void)b.combine_with_forward(br, fr.estimate, fr.residual);
    BidirectionalPushSolver bi(PushGraph::from_edges(2, {{0, 1, 1.0}}));
    (void)bi.solve_bidirectional(0, 1); (void)bi.adaptive_solve(0, 1);
    auto fp = forward_push_solve(m, {5.0, 4.0}, 1e-8, 1000);
    (void)fp.converged;
    auto sm = StreamingMatrix::from_triplets({{0, 0, 4.0}, {1, 1, 3.0}}, 2, 2, 1);
    size_t seen = 0;
    sm.multiply_vector_streaming({1.0, 1.0}, [&](size_t, const double *, size_t len) { seen += len; });
    (void)sm.memory_usage(); (void)sm.num_chunks();
    auto om = OptimizedSparseMatrix::from_triplets({{0, 0, 4.0}}, 1, 1);
    OptimizedConjugateGradientSolver cg;
    (void)cg.solve(om, {1.0});
    return (int)x.size();
}

int main(int argc, char **) { return use_everything(argc > 100); }
