"""Reference wire formats: parsing on CPU, replay through the B200 path on GPU."""
import json

import numpy as np
import pytest

import sublinear_b200 as sb
from sublinear_b200 import io as sio

DOC_3X3 = {"rows": 3, "cols": 3, "format": "dense", "data": [[4, -1, 0], [-1, 4, -1], [0, -1, 4]]}   # src/cli/index.ts:377-388


def coo_of(dense):
    a = np.asarray(dense, float)
    r, c = np.nonzero(a)
    return r.tolist(), c.tolist(), a[r, c].tolist()


def test_all_reference_spellings_parse_to_the_same_triplets():
    r, c, v = coo_of(DOC_3X3["data"])
    flat = {"rows": 3, "cols": 3, "format": "coo", "values": v, "rowIndices": r, "colIndices": c}            # types.ts:6-13
    nested = {"rows": 3, "cols": 3, "format": "coo", "data": {"values": v, "rowIndices": r, "colIndices": c}}  # solver.js:81-84
    fixture = {"matrix": DOC_3X3["data"], "size": 3, "rhs_vectors": {"ones": [1, 1, 1]}}
    mm = "%%MatrixMarket matrix coordinate real general\n% comment\n3 3 7\n" + \
         "\n".join(f"{i + 1} {j + 1} {x}" for i, j, x in zip(r, c, v)) + "\n"
    ref = sio.parse_matrix(DOC_3X3)
    for other in (sio.parse_matrix(flat), sio.parse_matrix(nested), sio.parse_matrix(fixture),
                  sio.parse_matrix(DOC_3X3["data"]), sio.parse_matrix_market(mm)):
        for a, b in zip(ref, other):
            assert np.array_equal(np.asarray(a), np.asarray(b))
    assert ref[3:] == (3, 3) and len(ref[2]) == 7


def test_format_errors():
    with pytest.raises(sb.SolverError) as e:
        sio.parse_matrix({"rows": 2, "cols": 2, "format": "csc", "values": []})
    assert e.value.variant == "UnsupportedMatrixFormat"
    with pytest.raises(sb.SolverError) as e:
        sio.parse_matrix({"rows": 2, "cols": 2, "format": "coo", "values": [1.0], "rowIndices": [0]})
    assert e.value.variant == "InvalidInput"
    with pytest.raises(sb.SolverError) as e:
        sio.parse_matrix({"rows": 5, "cols": 3, "format": "dense", "data": [[1, 0, 0], [0, 1, 0], [0, 0, 1]]})
    assert e.value.variant == "DimensionMismatch"


@pytest.mark.gpu
def test_replay_documented_examples(tmp_path, oracle):
    # the CLI help example and the MCP example (docs/reference/MCP_TOOL_TEST_RESULTS.md:49-55): true x = A^-1 b
    p = tmp_path / "m.json"
    p.write_text(json.dumps(DOC_3X3))
    m = sio.matrix_from_file(str(p))
    r = sb.NeumannSolver.new(100, 1e-12).solve(m, [1., 2., 1.], sb.SolverOptions(tolerance=1e-10))
    np.testing.assert_allclose(r.solution, np.linalg.solve(np.asarray(DOC_3X3["data"], float), [1., 2., 1.]), rtol=1e-9)
    mcp = {"rows": 3, "cols": 3, "format": "dense", "data": [[4, -1, 0], [-1, 4, -1], [0, -1, 3]]}
    r = sb.NeumannSolver.new(100, 1e-12).solve(sio.matrix_from_json(mcp), [1., 2., 1.], sb.SolverOptions(tolerance=1e-10))
    np.testing.assert_allclose(r.solution, [0.43902439, 0.75609756, 0.58536585], rtol=1e-7)   # not the sign-bugged [0.1463, ...]
