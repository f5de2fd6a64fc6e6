"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports exactly what include/sublinear_b200.h
declares, its host-only entry points (presets, partitioning, generator) behave like the reference, and compute entry
points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import sublinear_b200 as sb


def header_functions():
    src = open(sb.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(sb.LIB_PATH), "libsublinear_b200.so missing: run __graft_entry__.build()"
    L = C.CDLL(sb.LIB_PATH)
    declared = header_functions()
    assert len(declared) >= 50
    missing = [f for f in declared if not hasattr(L, f)]
    assert not missing, f"declared in the header but not exported: {missing}"
    out = subprocess.run(["nm", "-D", "--defined-only", sb.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (sb200_[a-z0-9_]+)", out)))
    undeclared = [f for f in exported if f not in declared]
    assert not undeclared, f"exported but not declared in the header: {undeclared}"
    sb.lib()  # the binding resolves every symbol it uses
    assert sb.lib().sb200_abi_version() == 1


def test_no_torch_types_or_link_dependency():
    out = subprocess.run(["ldd", sb.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "libnccl" not in out and "python" not in out
    hdr = re.sub(r"/\*.*?\*/", "", open(sb.HEADER_PATH).read(), flags=re.S)   # comments stripped
    assert "torch" not in hdr and "at::" not in hdr and "std::" not in hdr
    assert set(re.findall(r"#include <([^>]+)>", hdr)) == {"stddef.h", "stdint.h"}


def test_options_presets_match_reference():
    # SolverOptions::default / high_precision / fast / streaming (src/solver/mod.rs:47-116, tests :564-581)
    d = sb.SolverOptions.default()
    assert (d.tolerance, d.max_iterations, d.collect_stats, d.streaming_interval) == (1e-6, 1000, False, 0)
    assert (d.compute_error_bounds, d.error_bounds_tolerance, d.enable_profiling, d.random_seed) == (False, 1e-8, False, None)
    assert (d.convergence_mode, d.norm_type) == (0, 1)
    assert (d.mode, d.dominance, d.residual_check) == (sb.MODE_CORRECT, sb.DOMINANCE_ROW, sb.RESIDUAL_EVERY_5)
    h = sb.SolverOptions.high_precision()
    assert (h.tolerance, h.max_iterations, h.convergence_mode, h.collect_stats) == (1e-12, 5000, 4, True)
    assert (h.compute_error_bounds, h.error_bounds_tolerance) == (True, 1e-14)
    f = sb.SolverOptions.fast()
    assert (f.tolerance, f.max_iterations, f.error_bounds_tolerance) == (1e-3, 100, 1e-4)
    s = sb.SolverOptions.streaming(25)
    assert (s.tolerance, s.max_iterations, s.collect_stats, s.streaming_interval) == (1e-4, 1000, True, 25)
    assert (s.error_bounds_tolerance, s.enable_profiling) == (1e-6, True)


def test_neumann_solver_presets_match_reference():
    # src/solver/neumann.rs:48-80 and the creation test :563-573
    s = sb.NeumannSolver.new(16, 1e-8)
    assert s.config() == {"max_terms": 16, "series_tolerance": 1e-8, "adaptive_truncation": True, "cache_powers": True}
    assert sb.NeumannSolver.default().config()["max_terms"] == 50
    f = sb.NeumannSolver.fast()
    assert f.config() == {"max_terms": 20, "series_tolerance": 1e-6, "adaptive_truncation": False, "cache_powers": False}
    h = sb.NeumannSolver.high_precision()
    assert (h.config()["max_terms"], h.config()["series_tolerance"]) == (100, 1e-12)
    assert s.with_adaptive_truncation(False).with_power_caching(False).config()["adaptive_truncation"] is False
    assert s.algorithm_name() == "neumann"      # neumann.rs:464-466


def test_partition_rows_follows_reference_chunking():
    # chunk_size = ceil(rows / threads), contiguous (src/simd_ops.rs:219)
    assert [sb.partition_rows(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [sb.partition_rows(3, 8, r) for r in range(8)] == [(0, 1), (1, 2), (2, 3)] + [(3, 3)] * 5
    assert sb.partition_rows(10_000_000, 8, 7) == (8_750_000, 10_000_000)
    with pytest.raises(sb.SolverError):
        sb.partition_rows(10, 4, 4)


def test_library_generator_matches_oracle_generator(oracle):
    # two independent restatements of benches/performance_benchmarks.rs:12-43 must agree bit for bit
    for size, sparsity in [(1000, 0.02), (257, 0.5), (5000, 1e-4), (10, 0.0)]:
        rp, ci, v, b = sb.gen_bench_csr(size, sparsity)
        A, b2 = oracle.gen_bench_csr(size, sparsity)
        assert (A.row_ptr == rp).all() and (A.col_indices == ci).all() and (A.values == v).all() and (b == b2).all()
    rp, ci, v, b = sb.gen_bench_csr(1000, 0.02, 100, 900)
    A, b2 = oracle.gen_bench_csr(1000, 0.02, 100, 900)
    assert (A.row_ptr == rp).all() and (A.col_indices == ci).all() and (A.values == v).all() and (b == b2).all()


def test_host_side_validation_order_matches_from_triplets():
    # bounds / finiteness are checked on the host before any device work (src/matrix/mod.rs:166-187)
    with pytest.raises(sb.SolverError) as e:
        sb.SparseMatrix.from_triplets([2], [0], [1.0], 2, 2)
    assert e.value.variant == "IndexOutOfBounds"
    with pytest.raises(sb.SolverError) as e:
        sb.SparseMatrix.from_triplets([0], [2], [1.0], 2, 2)
    assert e.value.variant == "IndexOutOfBounds"
    with pytest.raises(sb.SolverError) as e:
        sb.SparseMatrix.from_triplets([0], [0], [float("inf")], 2, 2)
    assert e.value.variant == "InvalidInput"
    with pytest.raises(sb.SolverError) as e:
        sb.SparseMatrix.from_csr(np.array([0, 2, 1], np.uint32), [0, 1], [1.0, 2.0], 2, 2)
    assert e.value.variant == "InvalidSparseMatrix"


def test_compute_fails_loudly_without_gpu():
    if sb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sb.SolverError) as e:
        sb.SparseMatrix.from_triplets([0], [0], [1.0], 1, 1)
    assert e.value.variant == "AlgorithmError" and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """The product path must not import, link or read anything under oracle/ (or the reference)."""
    pkg = sb.PKG_DIR
    for root, _, files in os.walk(pkg):
        if "build" in root.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(root, f), errors="replace").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
                assert "/root/reference" not in txt, f
    out = subprocess.run(["ldd", sb.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def _build_cpp_unit_tests(tmp_path):
    exe = os.path.join(str(tmp_path), "test_reference_units")
    src = os.path.join(os.path.dirname(__file__), "cpp", "test_reference_units.cpp")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-o", exe, src, "-L", sb.PKG_DIR, "-lsublinear_b200",
                        f"-Wl,-rpath,{sb.PKG_DIR}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def test_cpp_mirror_of_reference_api_compiles_and_links(tmp_path):
    """cpp/sublinear.hpp (SparseMatrix / NeumannSolver / SolverOptions / SolverResult / SolverError over the C ABI)
    compiles with plain g++ and links against the library; the program is RUN by the gpu test below."""
    _build_cpp_unit_tests(tmp_path)


def test_cpp_mirror_covers_state_push_and_cg_api(tmp_path):
    """every wrapper of cpp/sublinear.hpp (stepping interface, push solvers, CG) compiles against the header and links
    against the library; nothing touches a device"""
    exe = os.path.join(str(tmp_path), "compile_api_mirror")
    src = os.path.join(os.path.dirname(__file__), "cpp", "compile_api_mirror.cpp")
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O0", "-Wall", "-o", exe, src, "-L", sb.PKG_DIR, "-lsublinear_b200",
                        f"-Wl,-rpath,{sb.PKG_DIR}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([exe], capture_output=True).returncode == 0


@pytest.mark.gpu
def test_cpp_reference_unit_tests_run(tmp_path):
    exe = _build_cpp_unit_tests(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "all passed" in r.stdout, r.stdout + r.stderr
