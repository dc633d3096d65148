"""TEST INFRASTRUCTURE: the reference's DokMatrix ingest semantics as a small pure-Python model (dict of dicts),
for checking the GPU ingest (and its emulation) on machines where the compiled reference is absent.
Follows io::readDokMatrix (src/runtime/IO.hpp:124-148), DokMatrix::set / explicitSymmetric
(src/runtime/SparseMatrix.hpp:204-208, 156-189) and CsrMatrix(const DokMatrix&) (:289-305).
Pinned against the compiled reference in tests/test_ingest_emu.py::test_model_matches_the_compiled_reference."""
import numpy as np


class NotSymmetric(Exception):
    pass


def dok_ingest(n, m, rows, cols, vals, symmetric, one_based=True, drop_upper=False):
    """Returns (row_ptr[n+1] with row_ptr[n] = true nnz, col_ind, values, nnzs_field)."""
    base = 1 if one_based else 0
    dok = {}
    nnzs = 0
    for i, j, v in zip(np.asarray(rows).tolist(), np.asarray(cols).tolist(), np.asarray(vals).tolist()):
        i -= base; j -= base
        if drop_upper and j > i:
            continue
        dok[(i, j)] = v       # DokMatrix::set: last value wins ...
        nnzs += 1             # ... and every call counts
    if symmetric:
        out, nnzs = {}, 0
        for (i, j), v in dok.items():
            out[(i, j)] = v
            nnzs += 1
            if i == j:
                continue
            if (j, i) in dok and dok[(j, i)] != v:
                raise NotSymmetric("Matrix is not symmetric")
            out[(j, i)] = v
            nnzs += 1
        dok = out
    elif drop_upper:
        nnzs = len(dok)
    keys = sorted(dok)
    counts = np.zeros(n + 1, np.int64)
    for i, _ in keys:
        counts[i + 1] += 1
    rp = np.cumsum(counts).astype(np.int32)
    ci = np.array([k[1] for k in keys], np.int32)
    va = np.array([dok[k] for k in keys], np.float64)
    return rp, ci, va, nnzs
