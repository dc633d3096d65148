"""Pair dictionaries of the pair-coded staged-ELL format (cask_b200/csrc/valuedict_logic.inl: BuildPairs) executed through
the host emulation of the device-logic backend (tests/emu).  A numpy model lays a matrix out exactly as plan_fill_kernel
does (entry (k, row = j*T + t) at (k*T + t)*4 + j, 16-bit x-cache positions, padding = value +0.0 at position 0); the
checks: decoding code -> (value, displacement + row) reproduces every stored (value, position) BIT FOR BIT, padding is
code 0, a stencil needs one pair per stencil point, more than 255 pairs in a slice raise the overflow flag.
CPU only; the GPU run is tests/test_gpu_x_value_dict.py."""
import numpy as np

import emu


def ell_slice(rows_cols_vals, slice_rows, x_pos):
    """Storage of ONE slice: rows_cols_vals[r] = list of (col, val) of local row r (ascending col); x_pos(col) = 16-bit
    x-cache position (>= 2).  Returns (width, vals, idx) in plan_fill_kernel's order."""
    T = slice_rows // 4
    width = max((len(r) for r in rows_cols_vals), default=0)
    vals = np.zeros(width * slice_rows)
    idx = np.zeros(width * slice_rows, np.uint16)
    for row, ent in enumerate(rows_cols_vals):
        t, j = row % T, row // T
        for k, (c, v) in enumerate(ent):
            e = (k * T + t) * 4 + j
            vals[e] = v
            idx[e] = x_pos(c)
    return width, vals, idx


def decode_and_check(val_off, widths, vals, idx, slice_rows, r):
    T = slice_rows // 4
    bits = vals.view(np.uint64)
    for q, (o, w) in enumerate(zip(val_off, widths)):
        n = r["npairs"][q]
        tab = r["table"][q]
        tv, td8 = tab["value_bits"], tab["disp8"].astype(np.int64)
        assert np.all(tab["zero"] == 0)
        if n:                                                  # slot 0: padding = +0.0 read slice_rows entries before the row
            assert tv[0] == 0 and td8[0] == -8 * slice_rows
        assert np.all(tv[n:] == 0) and np.all(td8[n:] == 0)
        assert np.all(td8 % 8 == 0)
        e = np.arange(w * slice_rows)
        rem = e % slice_rows
        row = (rem & 3) * T + (rem >> 2)
        codes = r["codes"][o:o + w * slice_rows].astype(np.int64)
        assert codes.max(initial=0) < max(n, 1)
        pos = td8[codes] // 8 + row                            # what the kernel computes, no special case for padding
        pad = idx[o:o + w * slice_rows] == 0
        assert np.array_equal(codes == 0, pad)
        assert np.array_equal(pos[~pad], idx[o:o + w * slice_rows].astype(np.int64)[~pad])
        assert np.all((pos[pad] >= -slice_rows) & (pos[pad] < 0))          # inside the zero region in front of the x cache
        assert np.array_equal(tv[codes], bits[o:o + w * slice_rows])      # padding value is +0.0 in both
        # pairs are distinct and listed in order of first appearance
        pairs = list(zip(tv[1:n].tolist(), td8[1:n].tolist()))
        assert len(set(pairs)) == len(pairs)
        first = {}
        for c in codes[codes > 0].tolist():
            first.setdefault(c, len(first) + 1)
        assert all(k == v for k, v in first.items())


def poisson2d_slices(N, slice_rows):
    """5-point stencil on an N x N grid, cut into slices of slice_rows rows; x cache of a slice = one contiguous window
    starting 2 positions in (the zero slots), so position = 2 + col - window_start."""
    n = N * N
    out = []
    for r0 in range(0, n, slice_rows):
        rows = []
        lo = max(0, r0 - N)
        for r in range(r0, min(n, r0 + slice_rows)):
            i, j = divmod(r, N)
            ent = []
            if i > 0: ent.append((r - N, -1.0))
            if j > 0: ent.append((r - 1, -1.0))
            ent.append((r, 4.0))
            if j < N - 1: ent.append((r + 1, -1.0))
            if i < N - 1: ent.append((r + N, -1.0))
            rows.append(ent)
        rows += [[] for _ in range(slice_rows - len(rows))]
        out.append(ell_slice(rows, slice_rows, lambda c, lo=lo: 2 + c - lo))
    return out


def concat(slices, slice_rows):
    widths = [s[0] for s in slices]
    off = np.concatenate([[0], np.cumsum(np.array(widths, np.int64) * slice_rows)])
    return off[:-1], widths, np.concatenate([s[1] for s in slices]), np.concatenate([s[2] for s in slices])


def test_five_point_stencil_needs_five_pairs_every_visiting_order():
    sr = 64
    val_off, widths, vals, idx = concat(poisson2d_slices(24, sr), sr)
    for order in (0, 1, 2):
        r = emu.pairdict(val_off, widths, vals, idx, sr, order)
        assert r["rc"] == 0 and not r["overflow"]
        assert r["max_pairs"] == 6                              # slot 0 + one pair per stencil point
        decode_and_check(val_off, widths, vals, idx, sr, r)
    assert emu.lib().emu_live_allocations() == 0


def test_random_slices_and_bit_patterns():
    rng = np.random.default_rng(2)
    sr = 32
    slices = []
    pool = np.array([1.5, -0.0, 0.0, np.inf, -2.25, 1e-300])
    nan = np.array([0x7ff8000000000001], np.uint64).view(np.float64)[0]
    for q in range(6):
        rows = []
        for r in range(sr if q != 3 else 0):
            cols = np.sort(rng.choice(40, rng.integers(0, 6), replace=False))
            rows.append([(int(c), float(pool[rng.integers(0, len(pool))]) if rng.random() < 0.9 else nan) for c in cols])
        rows += [[] for _ in range(sr - len(rows))]
        slices.append(ell_slice(rows, sr, lambda c: 2 + c))
    val_off, widths, vals, idx = concat(slices, sr)
    r = emu.pairdict(val_off, widths, vals, idx, sr)
    assert r["rc"] == 0 and not r["overflow"]
    decode_and_check(val_off, widths, vals, idx, sr, r)
    assert r["npairs"][3] == 0                                  # empty slice: nothing scanned


def test_255_pairs_fit_and_256_overflow():
    sr = 256
    # one entry per row, all rows the same column (displacement differs per row) -> sr distinct pairs
    def one_col(nrows):
        rows = [[(7, 3.0)] if r < nrows else [] for r in range(sr)]
        return ell_slice(rows, sr, lambda c: 2 + c)
    val_off, widths, vals, idx = concat([one_col(255)], sr)
    r = emu.pairdict(val_off, widths, vals, idx, sr)
    assert not r["overflow"] and r["npairs"][0] == 256 and r["max_pairs"] == 256
    decode_and_check(val_off, widths, vals, idx, sr, r)
    val_off, widths, vals, idx = concat([one_col(3), one_col(256), one_col(2)], sr)
    r = emu.pairdict(val_off, widths, vals, idx, sr)
    assert r["overflow"] and r["npairs"].tolist() == [4, 0, 3]   # the caller drops the whole plan to the next format
    assert emu.pairdict([], [], np.zeros(0), np.zeros(0, np.uint16))["rc"] == 0
    assert emu.lib().emu_live_allocations() == 0


def plan_model(n, rp, ci, va, pipes, sr=1024):
    """numpy model of plan.cu for matrices whose slices are all staged: row stripes (Spmv.cpp:334-364) cut into slices of
    sr rows, x-cache position = 2 + 16 * (rank of the column's granule among the slice's referenced granules) + col % 16,
    ELL storage order of plan_fill_kernel."""
    T = sr // 4
    slices, start = [], 0
    for q in range(pipes):
        rows = (n - start) if q == pipes - 1 else n // pipes
        slices += [(start + r, min(sr, rows - r)) for r in range(0, rows, sr)]
        start += rows
    widths, vals, idxs = [], [], []
    for r0, nr in slices:
        gran = np.unique(ci[rp[r0]:rp[r0 + nr]] >> 4)
        rank = {g: i for i, g in enumerate(gran.tolist())}
        lens = np.diff(rp[r0:r0 + nr + 1])
        w = int(lens.max()) if nr else 0
        v, ix = np.zeros(w * sr), np.zeros(w * sr, np.uint16)
        for row in range(nr):
            t, j = row % T, row // T
            for k in range(lens[row]):
                c = int(ci[rp[r0 + row] + k])
                e = (k * T + t) * 4 + j
                v[e], ix[e] = va[rp[r0 + row] + k], 2 + rank[c >> 4] * 16 + (c & 15)
        widths.append(w); vals.append(v); idxs.append(ix)
    off = np.concatenate([[0], np.cumsum(np.array(widths, np.int64) * sr)])
    return off[:-1], widths, np.concatenate(vals), np.concatenate(idxs)


def test_one_pair_per_stencil_point_on_the_plan_layout(oracle):
    """BASELINE's stencils, down-scaled, laid out as the GPU partitioner lays them out (granule-ranked x windows, two
    stripes): every slice needs exactly one pair per stencil point - what tests/test_gpu_x_value_dict.py asserts of the
    GPU plan - and decoding reproduces every stored (value, position)."""
    for gen, N, points in (("gen_poisson2d", 96, 5), ("gen_poisson3d27", 14, 27), ("gen_convdiff3d7", 16, 7)):
        n, rp, ci, va = getattr(oracle, gen)(N)
        for pipes in (1, 2):
            off, widths, vals, idx = plan_model(n, rp, ci, va, pipes)
            r = emu.pairdict(off, widths, vals, idx, 1024)
            assert r["rc"] == 0 and not r["overflow"] and r["max_pairs"] == points + 1, (gen, pipes, r["max_pairs"])
            decode_and_check(off, widths, vals, idx, 1024, r)


def test_decoded_product_is_bit_identical_to_the_reference_dot(oracle):
    """The consumer loop of the pair-coded kernel, restated in numpy on the modelled plan: x cache = 1024 zeros, two zero
    slots, then the referenced 16-double granules of x; per row, ELL columns in ascending order, product then add, the x
    entry found at displacement + row with NO special case for padding.  y equals CsrMatrix::dot bit for bit - also when x
    holds Inf and NaN, which a padding entry multiplied by a real x entry would spread to rows that never reference them."""
    sr = 1024
    T = sr // 4
    n, rp, ci, va = oracle.gen_poisson2d(40)
    x = np.random.default_rng(4).standard_normal(n)
    x[[5, 900, 1599]] = [np.inf, np.nan, -np.inf]
    exp = oracle.csr_dot(n, rp, ci, va, x)
    off, widths, vals, idx = plan_model(n, rp, ci, va, 1, sr)
    r = emu.pairdict(off, widths, vals, idx, sr)
    assert r["rc"] == 0 and not r["overflow"]
    y = np.zeros(n)
    with np.errstate(invalid="ignore"):
        for q, (o, w) in enumerate(zip(off, widths)):
            r0 = q * sr
            nr = min(sr, n - r0)
            gran = np.unique(ci[rp[r0]:rp[r0 + nr]] >> 4)
            xs = np.zeros(sr + 2 + 16 * len(gran))                 # [zero region | zero slots | granules]
            for i, g in enumerate(gran.tolist()):
                seg = x[g * 16:min(n, g * 16 + 16)]
                xs[sr + 2 + 16 * i: sr + 2 + 16 * i + len(seg)] = seg
            tab = r["table"][q]
            tv = tab["value_bits"].view(np.float64)
            td = tab["disp8"].astype(np.int64) // 8
            codes = r["codes"][o:o + w * sr].astype(np.int64)
            for row in range(nr):
                t, j = row % T, row // T
                acc = 0.0
                for k in range(w):
                    c = codes[(k * T + t) * 4 + j]
                    acc = acc + tv[c] * xs[sr + td[c] + row]
                y[r0 + row] = acc
    assert np.array_equal(np.isnan(y), np.isnan(exp))
    ok = ~np.isnan(exp)
    assert np.array_equal(y[ok].view(np.uint64), exp[ok].view(np.uint64))
