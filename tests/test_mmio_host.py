"""Matrix Market tokeniser of libcask_b200.so (cask_b200/csrc/mmio.cpp; host code, no GPU needed) against
io::readHeader / readDokMatrix / readVector of the reference (src/runtime/IO.hpp:60-148): live against the compiled
reference where /root/reference exists, and against the committed golden CSRs everywhere."""
import glob
import os

import numpy as np
import pytest

import cask_b200 as cb
from oracle import refbind as R

REF = "/root/reference"
needs_ref = pytest.mark.skipif(not (R.available() and os.path.isdir(REF)), reason="needs /root/reference")


def write_mtx(path, n, m, rows, cols, vals, symmetry="general", header=None, sep="\n"):
    with open(path, "w") as f:
        f.write((header or "%%MatrixMarket matrix coordinate real " + symmetry) + "\n")
        f.write("% written by the test-suite\n%\n")
        f.write("%d %d %d\n" % (n, m, len(vals)))
        f.write(sep.join("%d %d %s" % (r, c, repr(float(v))) for r, c, v in zip(rows, cols, vals)))
        f.write("\n")


def coo_of(n, rp, ci, va):
    rows = np.repeat(np.arange(n), np.diff(rp)) + 1
    return rows.astype(np.int32), (ci + 1).astype(np.int32), va


def test_round_trip_of_a_golden_matrix(tmp_path, golden):
    n, m, rp, ci, va = golden.csr("test_cage6")
    rows, cols, vals = coo_of(n, rp, ci, va)
    p = str(tmp_path / "a.mtx")
    write_mtx(p, n, m, rows, cols, vals)
    info, r, c, v = cb.mm_read_coo(p)
    assert info == {"type": "matrix", "format": "coordinate", "data_type": "real", "symmetry": "general", "n": n, "m": m,
                    "entries": len(va)}
    assert np.array_equal(r, rows) and np.array_equal(c, cols) and np.array_equal(v, vals)


def test_entries_are_tokens_not_lines(tmp_path):
    """operator>> at IO.hpp:143 reads whitespace-separated tokens: entries may share or straddle lines."""
    p = str(tmp_path / "t.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n3 3 4\n1 1 1.5 2\n2\n\t-2e0   3 1 +4 3 3 .25 9 9 9\n")
    info, r, c, v = cb.mm_read_coo(p)
    assert r.tolist() == [1, 2, 3, 3] and c.tolist() == [1, 2, 1, 3] and v.tolist() == [1.5, -2.0, 4.0, 0.25]


@pytest.mark.parametrize("header", [
    "%%MatrixMarket matrix coordinate pattern general",      # pattern / complex are not in the reference's regex
    "%%MatrixMarket matrix coordinate complex general",
    "%%MatrixMarket matrix coordinate real skew-symmetric",
    "%%MatrixMarket matrix coordinate real general ",        # regex_match: nothing may follow
    "%%MatrixMarket  matrix coordinate real general",        # single spaces only
    "%%MatrixMarket matrix coordinate real general\r",       # CRLF files are rejected by the reference too
    "%MatrixMarket matrix coordinate real general",
    "%%matrixmarket matrix coordinate real general",
])
def test_headers_the_reference_rejects(tmp_path, header):
    p = str(tmp_path / "h.mtx")
    write_mtx(p, 2, 2, [1], [1], [1.0], header=header)
    with pytest.raises(cb.CaskError) as e:
        cb.mm_read_info(p)
    assert e.value.code == cb.ERR_INVALID_ARGUMENT and "Not a valid MatrixMarket file in " + p in e.value.message
    if R.available():
        with pytest.raises(RuntimeError) as e2:
            R.RefMatrix.read(p)
        assert "Not a valid MatrixMarket file" in str(e2.value)


def test_missing_file_message():
    with pytest.raises(cb.CaskError) as e:
        cb.mm_read_info("/nonexistent/x.mtx")
    assert e.value.message == "File not found /nonexistent/x.mtx"  # IO.hpp:63-64


def test_malformed_and_truncated_files_fail_loudly(tmp_path):
    p = str(tmp_path / "m.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n2 x 3.0\n")
    with pytest.raises(cb.CaskError) as e:
        cb.mm_read_coo(p)
    assert "Malformed entry 2" in e.value.message
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n2 2 3\n1 1 1.0\n2 2 3.0\n")
    with pytest.raises(cb.CaskError) as e:
        cb.mm_read_coo(p)
    assert "ends after 2 of 3 entries" in e.value.message
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% only comments\n")
    with pytest.raises(cb.CaskError):
        cb.mm_read_info(p)
    # a size line that claims 2^31-1 entries in a 70-byte file is an error code, not a 32 GB allocation / std::terminate
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n2 2 2147483647\n1 1 1.0\n")
    with pytest.raises(cb.CaskError) as e:
        cb.mm_read_coo(p)
    assert "ends after 1 of 2147483647 entries" in e.value.message
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n2147483647 1\n1.0\n")
    with pytest.raises(cb.CaskError) as e:
        cb.mm_read_vector(p)
    assert "shorter than" in e.value.message


def test_large_file_takes_the_multithreaded_path(tmp_path):
    rng = np.random.default_rng(0)
    L = 200_000  # ~4.5 MB of text: split over the host cores at token boundaries
    rows = rng.integers(1, 50_001, L).astype(np.int32)
    cols = rng.integers(1, 50_001, L).astype(np.int32)
    vals = rng.standard_normal(L)
    p = str(tmp_path / "big.mtx")
    write_mtx(p, 50_000, 50_000, rows, cols, vals, sep=" ")  # one enormous line: chunks cannot rely on newlines
    info, r, c, v = cb.mm_read_coo(p)
    assert os.path.getsize(p) > (1 << 20)
    assert np.array_equal(r, rows) and np.array_equal(c, cols) and np.array_equal(v, vals)


def test_vectors(tmp_path):
    p = str(tmp_path / "v.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix array real general\n% c\n4 1\n1\n2.5\n-3\n4e-1\n")
    assert cb.mm_read_vector(p).tolist() == [1.0, 2.5, -3.0, 0.4]
    with open(p, "w") as f:  # coordinate vectors are NOT rebased by the reference (v[a] = val, IO.hpp:99)
        f.write("%%MatrixMarket matrix coordinate real general\n4 1 2\n1 1 7\n3 1 9\n")
    assert cb.mm_read_vector(p).tolist() == [0.0, 7.0, 0.0, 9.0]
    if R.available():
        assert R.read_vector(p).tolist() == [0.0, 7.0, 0.0, 9.0]


@needs_ref
def test_every_reference_file_reads_like_the_reference():
    """All of test/matrices, test/systems and test/test-benchmark: same header verdict; vectors equal; for matrices
    the entry list, pushed through the reference's own DokMatrix semantics in numpy, reproduces io::readMatrix."""
    files = sorted(glob.glob(REF + "/test/matrices/*.mtx") + glob.glob(REF + "/test/systems/*.mtx") +
                   glob.glob(REF + "/test/test-benchmark/*.mtx"))
    assert len(files) >= 40
    for f in files:
        info = cb.mm_read_info(f)
        if info["format"] != "coordinate":
            assert np.array_equal(cb.mm_read_vector(f), R.read_vector(f))
            continue
        _, r, c, v = cb.mm_read_coo(f)
        ref = R.RefMatrix.read(f, sym_lower=False)
        rp, ci, va = ref.csr()
        d = {}
        for i, j, x in zip(r.tolist(), c.tolist(), v.tolist()):
            d[(i - 1, j - 1)] = x
            if info["symmetry"] == "symmetric":
                d[(j - 1, i - 1)] = x
        keys = sorted(d)
        assert len(keys) == len(va), f
        assert np.array_equal(np.array([k[1] for k in keys], np.int32), ci), f
        assert np.array_equal(np.array([d[k] for k in keys]), va), f
