"""Access-pattern model of the C3 (R-MAT) x gather, CPU only - what a GPU capture should be read against.

    python profiles/c3_gather_model.py [scale ...] > profiles/rN_c3_gather_model.md

For R-MAT matrices from the oracle's generator ((a, b, c, d) = (0.57, 0.19, 0.19, 0.05), edge factor 15, duplicates merged,
the BASELINE configs[2] recipe at smaller scales) it counts
  * the share of all gathers that go to the K most popular columns (what a K-entry shared-memory table would serve);
  * distinct 128-byte lines and 32-byte sectors among 32 consecutive nonzeros in CSR order - the L1 tag-stage wavefronts and
    the L2 sectors of one warp-wide gather in the gather-CSR kernels (both variants walk the nonzeros in this order);
  * the lines left per warp-wide gather when the K = 15 360 most popular columns come from shared memory.
The trend over the scale is what extrapolates to scale 25 (the share of a fixed-size table falls by ~0.7x per two scales)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from oracle import oraclebind as O
    O.build()
    scales = [int(a) for a in sys.argv[1:]] or [18, 20, 22]
    print("# C3 gather model (CPU simulation of the access pattern; R-MAT from oracle/cask_oracle.c)\n")
    print("| scale | rows | nnz | top 4 096 cols | top 15 360 | top 24 576 | lines / 32 nnz | sectors / 32 nnz | lines / 32 nnz with a 15 360-entry table |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
    for scale in scales:
        n, rp, ci, va = O.gen_rmat(scale, 15, 1)
        nnz = len(ci)
        deg = np.bincount(ci, minlength=n)
        order = np.argsort(-deg, kind="stable")
        cum = np.cumsum(deg[order]) / nnz
        m = min((nnz // 32) * 32, 32 * 200000)
        cols = ci[:m]
        lines = np.sort((cols >> 4).reshape(-1, 32), axis=1)
        sect = np.sort((cols >> 2).reshape(-1, 32), axis=1)
        nl = ((np.diff(lines, axis=1) != 0).sum(1) + 1).mean()
        ns = ((np.diff(sect, axis=1) != 0).sum(1) + 1).mean()
        hot = np.zeros(n, bool)
        hot[order[:15360]] = True
        l2 = np.sort(np.where(~hot[cols].reshape(-1, 32), (cols >> 4).reshape(-1, 32), -1), axis=1)
        nh = ((np.diff(l2, axis=1) != 0).sum(1) + 1 - (l2[:, 0] == -1)).mean()
        print("| %d | %d | %d | %.1f %% | %.1f %% | %.1f %% | %.1f | %.1f | %.1f |" % (
            scale, n, nnz, 100 * cum[4095], 100 * cum[15359], 100 * cum[24575], nl, ns, nh))


if __name__ == "__main__":
    main()
