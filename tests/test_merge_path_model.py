"""Model of the merge-path gather kernel (cask_b200/csrc/spmv.cu: spmv_csr_merge_kernel + spmv_csr_merge_fixup_kernel,
plan.cu: merge_tiles_kernel) in plain Python, statement for statement, with small tile parameters: the tile split rule,
every thread's walk over the merge path, the segmented scan of the partial sums across "threads" and "warps", the
per-tile carries and the fix-up chain.  Checked against the CSR row sums on matrices with empty rows, rows longer than
several tiles and runs that start in the middle of the matrix.  The CUDA code itself is checked on the GPU
(tests/test_gpu_spmv.py::test_merge_path_*); this file pins the ALGORITHM on CPU."""
import numpy as np
import pytest


def split(row_ptr, ra, rb, ka, kb, diag):
    nrows, nnz = rb - ra, kb - ka
    total = nrows + nnz
    diag = min(diag, total)
    lo, hi = max(diag - nnz, 0), min(diag, nrows)
    while lo < hi:
        mid = (lo + hi) >> 1
        if row_ptr[ra + mid + 1] - ka <= diag - mid - 1:
            lo = mid + 1
        else:
            hi = mid
    return ra + lo, ka + (diag - lo)


def build_tiles(row_ptr, runs, tile):
    tiles = []
    for ra, rb in runs:
        ka, kb = row_ptr[ra], row_ptr[rb]
        total = (rb - ra) + (kb - ka)
        for i in range((total + tile - 1) // tile):
            r0, k0 = split(row_ptr, ra, rb, ka, kb, i * tile)
            r1, k1 = split(row_ptr, ra, rb, ka, kb, (i + 1) * tile)
            tiles.append((r0, k0, r1, k1))
    return tiles


def run_tile(t, n_rows, row_ptr, col, val, x, y, threads, items, warp):
    r0, k0, r1, k1 = t
    nnzT, nrowsT = k1 - k0, r1 - r0
    rend = [row_ptr[r0 + r + 1] - k0 if r0 + r < n_rows else 0x7fffffff for r in range(nrowsT + 1)]
    prod = [val[k0 + j] * x[col[k0 + j]] for j in range(nnzT)]
    total = nnzT + nrowsT
    key, sv, r_start, has_first, first_val = [0] * threads, [0.0] * threads, [0] * threads, [False] * threads, [0.0] * threads
    for tid in range(threads):
        diag = min(tid * items, total)
        diag_end = min(diag + items, total)
        lo, hi = max(0, diag - nnzT), min(diag, nrowsT)
        while lo < hi:
            mid = (lo + hi) >> 1
            if rend[mid] <= diag - mid - 1:
                lo = mid + 1
            else:
                hi = mid
        r, k = lo, diag - lo
        r_start[tid] = r
        acc = 0.0
        for _ in range(diag, diag_end):
            if k < rend[r]:
                acc += prod[k]
                k += 1
            else:
                if not has_first[tid]:
                    has_first[tid], first_val[tid] = True, acc
                else:
                    y[r0 + r] = acc
                acc = 0.0
                r += 1
        key[tid], sv[tid] = r, acc
    assert all(key[i] <= key[i + 1] for i in range(threads - 1))
    # warp-level segmented scan (Hillis-Steele on key equality)
    nw = threads // warp
    d = 1
    while d < warp:
        nk, nv = list(key), list(sv)
        for tid in range(threads):
            lane = tid % warp
            if lane >= d and key[tid - d] == key[tid]:
                nv[tid] = sv[tid] + sv[tid - d]
        sv = nv
        d <<= 1
    tail_key = [key[w * warp + warp - 1] for w in range(nw)]
    tail_val = [sv[w * warp + warp - 1] for w in range(nw)]
    head_key = [key[w * warp] for w in range(nw)]
    pre_key, pre_val = [-1] * nw, [0.0] * nw
    for w in range(1, nw):
        pv = tail_val[w - 1]
        if head_key[w - 1] == tail_key[w - 1] and pre_key[w - 1] == tail_key[w - 1]:
            pv += pre_val[w - 1]
        pre_key[w], pre_val[w] = tail_key[w - 1], pv
    for tid in range(threads):
        if key[tid] == pre_key[tid // warp]:
            sv[tid] += pre_val[tid // warp]
    for tid in range(threads):
        if tid % warp == 0:
            pk, pv = pre_key[tid // warp], pre_val[tid // warp]
        else:
            pk, pv = key[tid - 1], sv[tid - 1]
        if has_first[tid]:
            y[r0 + r_start[tid]] = pv + first_val[tid] if pk == r_start[tid] else first_val[tid]
    return sv[threads - 1] if key[threads - 1] == nrowsT else 0.0


def fixup(tiles, row_ptr, carry, y):
    for i, t in enumerate(tiles):
        open_ = t[3] > row_ptr[t[2]]
        head = open_
        if open_ and i > 0:
            q = tiles[i - 1]
            head = not (q[2] == t[2] and q[3] > row_ptr[q[2]] and q[3] == t[1])
        if head:
            s = carry[i]
            for j in range(i + 1, len(tiles)):
                u = tiles[j]
                if u[2] != t[2] or u[1] != tiles[j - 1][3]:
                    break
                s += carry[j]
            y[t[2]] = s + y[t[2]]


def model_spmv(n, row_ptr, col, val, x, runs, threads=8, items=3, warp=4):
    tiles = build_tiles(row_ptr, runs, threads * items)
    y = np.full(n, np.nan)
    carry = [run_tile(t, n, row_ptr, col, val, x, y, threads, items, warp) for t in tiles]
    fixup(tiles, row_ptr, carry, y)
    return y, tiles


def random_csr(rng, n, m, lens):
    row_ptr = np.zeros(n + 1, np.int64)
    row_ptr[1:] = np.cumsum(lens)
    col = np.concatenate([np.sort(rng.choice(m, k, replace=False)) for k in lens]) if row_ptr[-1] else np.zeros(0, np.int64)
    val = rng.standard_normal(int(row_ptr[-1]))
    return row_ptr, col.astype(np.int64), val


@pytest.mark.parametrize("seed", range(12))
def test_model_matches_csr_row_sums(seed):
    rng = np.random.default_rng(seed)
    n, m = int(rng.integers(1, 120)), 400
    lens = rng.choice([0, 0, 0, 1, 2, 3, 5, 9, 40, 130, 300], size=n)
    row_ptr, col, val = random_csr(rng, n, m, lens)
    x = rng.standard_normal(m)
    exp = np.array([np.dot(val[row_ptr[i]:row_ptr[i + 1]], x[col[row_ptr[i]:row_ptr[i + 1]]]) for i in range(n)])
    threads, items, warp = [(8, 3, 4), (16, 5, 4), (4, 1, 2), (32, 17, 8)][seed % 4]
    y, tiles = model_spmv(n, row_ptr, col, val, x, [(0, n)], threads, items, warp)
    assert not np.isnan(y).any()
    scale = np.array([np.abs(val[row_ptr[i]:row_ptr[i + 1]] * x[col[row_ptr[i]:row_ptr[i + 1]]]).sum() for i in range(n)])
    assert np.all(np.abs(y - exp) <= 1e-12 * np.maximum(scale, 1e-300) + 0.0)
    # the tiles partition the merge path exactly
    assert tiles[0][:2] == (0, 0) and tiles[-1][2:] == (n, row_ptr[n])
    assert all(a[2:] == b[:2] for a, b in zip(tiles, tiles[1:]))


def test_model_runs_in_the_middle_and_all_empty_rows():
    rng = np.random.default_rng(99)
    n, m = 90, 64
    lens = rng.choice([0, 1, 4, 60], size=n)
    lens[40:70] = 0                       # a stretch of empty rows: tiles made of row ends only
    row_ptr, col, val = random_csr(rng, n, m, lens)
    x = rng.standard_normal(m)
    runs = [(10, 35), (35, 80)]           # two adjacent runs (the interior / halo split) starting mid-matrix
    y, tiles = model_spmv(n, row_ptr, col, val, x, runs, 8, 3, 4)
    exp = np.array([np.dot(val[row_ptr[i]:row_ptr[i + 1]], x[col[row_ptr[i]:row_ptr[i + 1]]]) for i in range(n)])
    assert np.isnan(y[:10]).all() and np.isnan(y[80:]).all()
    assert np.allclose(y[10:80], exp[10:80], rtol=0, atol=1e-12)
    assert (y[40:70] == 0).all()


def test_one_row_spanning_many_tiles():
    rng = np.random.default_rng(5)
    lens = np.array([2, 0, 500, 1, 0, 0, 3])
    row_ptr, col, val = random_csr(rng, len(lens), 600, lens)
    x = rng.standard_normal(600)
    y, tiles = model_spmv(len(lens), row_ptr, col, val, x, [(0, len(lens))], 8, 3, 4)
    exp = np.array([np.dot(val[row_ptr[i]:row_ptr[i + 1]], x[col[row_ptr[i]:row_ptr[i + 1]]]) for i in range(len(lens))])
    assert len(tiles) > 15 and np.allclose(y, exp, rtol=0, atol=1e-11)
