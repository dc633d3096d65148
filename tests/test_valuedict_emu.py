"""Value dictionaries of the coded staged-ELL format (cask_b200/csrc/valuedict_logic.inl) executed through the host
emulation of the device-logic backend (tests/emu): table[code] reproduces every stored value BIT FOR BIT (signed
zeros and NaN payloads included), tables list distinct bit patterns in order of first appearance, a slice with more
than 256 distinct values raises the overflow flag.  CPU only; the GPU run is tests/test_gpu_x_value_dict.py."""
import numpy as np

import emu


def layout(widths, slice_rows):
    off = np.concatenate([[0], np.cumsum(np.array(widths, np.int64) * slice_rows)])
    return off[:-1], int(off[-1])


def check(val_off, widths, vals, slice_rows, r):
    bits = vals.view(np.uint64)
    for q, (o, w) in enumerate(zip(val_off, widths)):
        seg = bits[o:o + w * slice_rows]
        n = r["ndict"][q]
        tab = r["table"][q].view(np.uint64)
        _, first = np.unique(seg, return_index=True)
        order = seg[np.sort(first)]                       # distinct patterns in order of first appearance
        assert n == len(order) and np.array_equal(tab[:n], order)
        assert np.all(tab[n:] == 0)
        codes = r["codes"][o:o + w * slice_rows]
        assert codes.max(initial=0) < max(n, 1)
        assert np.array_equal(tab[codes], seg)            # the looked-up double is the stored double, bit for bit


def test_stencil_like_slices_every_visiting_order():
    rng = np.random.default_rng(0)
    widths = [5, 27, 7, 1, 0, 3]
    sr = 64
    val_off, total = layout(widths, sr)
    pool = np.array([4.0, -1.0, 0.0, 26.0, -0.0, 6.0, 1e-300, np.inf])
    vals = pool[rng.integers(0, len(pool), total)]
    for order in (0, 1, 2):
        r = emu.valuedict(val_off, widths, vals, sr, order)
        assert r["rc"] == 0 and not r["overflow"] and r["max_entries"] <= len(pool)
        check(val_off, widths, vals, sr, r)
    assert r["ndict"][4] == 0                                 # empty slice: empty table


def test_bit_patterns_not_values():
    nan1 = np.array([0x7ff8000000000001], np.uint64).view(np.float64)[0]
    nan2 = np.array([0x7ff8000000000002], np.uint64).view(np.float64)[0]
    vals = np.array([0.0, -0.0, nan1, nan2, 0.0, nan1, -0.0, 1.0])
    r = emu.valuedict([0], [2], vals, 4)
    assert r["ndict"][0] == 5 and r["codes"].tolist() == [0, 1, 2, 3, 0, 2, 1, 4]
    check([0], [2], vals, 4, r)


def test_exactly_256_fits_and_257_overflows():
    sr = 128
    vals = np.arange(256, dtype=np.float64).repeat(2)[: 4 * sr]
    vals = np.concatenate([vals, np.arange(256, dtype=np.float64)[::-1]])      # 256 distinct values, 6 columns
    r = emu.valuedict([0], [6], vals, sr)
    assert not r["overflow"] and r["ndict"][0] == 256 and r["max_entries"] == 256
    check([0], [6], vals, sr, r)
    vals2 = np.concatenate([np.full(sr, 7.0), np.arange(3 * sr, dtype=np.float64) + 100.0, np.full(sr, 7.0)])
    widths = [1, 3, 1]
    val_off, _ = layout(widths, sr)
    r = emu.valuedict(val_off, widths, vals2, sr)
    assert r["overflow"] and r["ndict"].tolist() == [1, 0, 1]   # the other slices are still coded; the caller drops all


def test_random_values_overflow_and_no_leak():
    vals = np.random.default_rng(1).standard_normal(3 * 1024)
    r = emu.valuedict([0], [3], vals)
    assert r["rc"] == 0 and r["overflow"] and r["max_entries"] == 0
    r = emu.valuedict([], [], np.zeros(0))
    assert r["rc"] == 0 and not r["overflow"]
    assert emu.lib().emu_live_allocations() == 0
