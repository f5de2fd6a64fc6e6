"""Wire / disk formats either side of the path: the reference's JSON matrix schemas and its
Matrix Market reader, so that reference inputs can be replayed through the B200 path. Parsing is host-side Python and
returns COO triplets (rows, cols, vals, nrows, ncols); building the device matrix goes through the C ABI as usual.

Accepted (all from the reference):
  * dense      {"rows","cols","format":"dense","data":[[...],...]}          (src/core/types.ts:15-20, src/cli/index.ts:377-388)
  * COO, flat  {"rows","cols","format":"coo","values","rowIndices","colIndices"}   (src/core/types.ts:6-13)
  * COO, nested {"rows","cols","format":"coo","data":{"values","rowIndices","colIndices"}}  (src/solver.js:81-84, bin/cli.js:484-490)
  * fixture    {"matrix":[[...]],"size",...,"rhs_vectors":{...}}           (scripts/linear_systems/test_matrices/*.json)
  * bare array [[...],...]                                                   (bin/cli.js:455-461)
  * Matrix Market coordinate text                                            (bin/cli.js:463-491: 1-based, whitespace separated)
"""
from __future__ import annotations

import json

import numpy as np

from . import SolverError, SparseMatrix


def _dense_to_coo(a):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2:
        raise SolverError(4, "dense matrix data must be an array of rows")
    r, c = np.nonzero(a)          # row-major order, zeros filtered: SparseMatrix::from_dense (src/matrix/mod.rs:202-223)
    return r.astype(np.uint64), c.astype(np.uint64), a[r, c], a.shape[0], a.shape[1]


def parse_matrix(obj):
    """JSON object / array of the reference -> (rows, cols, vals, nrows, ncols)."""
    if isinstance(obj, (list, tuple, np.ndarray)):
        return _dense_to_coo(obj)
    if not isinstance(obj, dict):
        raise SolverError(4, "matrix must be a JSON object or an array of rows")
    if "matrix" in obj and "format" not in obj:          # linear_systems fixture
        return _dense_to_coo(obj["matrix"])
    fmt = obj.get("format", "dense")
    if fmt == "dense":
        r, c, v, nr, nc = _dense_to_coo(obj["data"])
        if "rows" in obj and (int(obj["rows"]) != nr or int(obj.get("cols", nc)) != nc):
            raise SolverError(5, f"dense data is {nr}x{nc}, header says {obj['rows']}x{obj.get('cols')}")
        return r, c, v, nr, nc
    if fmt == "coo":
        src = obj["data"] if isinstance(obj.get("data"), dict) else obj      # nested or flat spelling
        try:
            v = np.asarray(src["values"], dtype=np.float64)
            r = np.asarray(src["rowIndices"], dtype=np.int64)
            c = np.asarray(src["colIndices"], dtype=np.int64)
        except KeyError as e:
            raise SolverError(4, f"COO matrix lacks {e}") from None
        if not (len(v) == len(r) == len(c)):
            raise SolverError(5, "COO arrays differ in length")
        if (r < 0).any() or (c < 0).any():
            raise SolverError(8, "negative index in COO matrix")
        return r.astype(np.uint64), c.astype(np.uint64), v, int(obj["rows"]), int(obj["cols"])
    raise SolverError(6, f"unsupported matrix format {fmt!r}")


def parse_matrix_market(text: str):
    """bin/cli.js:463-491 — comment lines start with '%', then 'rows cols entries', then 1-based 'row col value'."""
    lines = text.strip().split("\n")
    h = 0
    while lines[h].startswith("%"):
        h += 1
    nr, nc, _ = (int(float(t)) for t in lines[h].split()[:3])
    r, c, v = [], [], []
    for ln in lines[h + 1:]:
        if ln.strip():
            a, b, x = ln.split()[:3]
            r.append(int(a) - 1)
            c.append(int(b) - 1)
            v.append(float(x))
    return np.asarray(r, np.uint64), np.asarray(c, np.uint64), np.asarray(v, np.float64), nr, nc


def load_matrix_file(path: str):
    with open(path) as f:
        text = f.read()
    if path.endswith((".mtx", ".mm")) or text.lstrip().startswith("%%MatrixMarket"):
        return parse_matrix_market(text)
    return parse_matrix(json.loads(text))


def matrix_from_json(obj) -> SparseMatrix:
    r, c, v, nr, nc = parse_matrix(obj)
    return SparseMatrix.from_triplets(r, c, v, nr, nc)


def matrix_from_file(path: str) -> SparseMatrix:
    r, c, v, nr, nc = load_matrix_file(path)
    return SparseMatrix.from_triplets(r, c, v, nr, nc)
