"""Host half of the Matrix Market ingest timed beside the reference's own reader, CPU only.

    python profiles/ingest_host_timing.py [file.mtx ...] > profiles/rN_ingest_host_timing.md

reference  io::readMatrix (src/runtime/IO.hpp:151-163) compiled in place (oracle/_ref): parse + DokMatrix hash-map
           build + CsrMatrix conversion, single thread - the whole ingest of the reference; best of 2.
tokeniser  cask_b200_mm_read_coo (cask_b200/csrc/mmio.cpp): the text -> (row, col, value) arrays on all host cores,
           best of 5.  The dictionary-of-keys semantics that follow (sort, last-value-wins, symmetric expansion) run on
           the GPU (csrc/ingest.cu) and are not part of this CPU-only figure.
Needs /root/reference (the files and the compiled reader); it is a development-container measurement, not a test."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = ["/root/reference/test/test-benchmark/psmigr_2.mtx", "/root/reference/test/matrices/TSOPF_RS_b39_c7.mtx",
           "/root/reference/test/matrices/OPF_6000.mtx", "/root/reference/test/test-benchmark/raefsky1.mtx"]


def main():
    import cask_b200 as cb
    from oracle import refbind as R
    files = sys.argv[1:] or DEFAULT
    print("# Matrix Market ingest, host side: reference reader vs the multi-threaded tokeniser (%d host threads)\n" % os.cpu_count())
    print("| file | bytes | entries | reference `io::readMatrix` (s) | tokeniser (s) | tokeniser MB/s | ratio |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for f in files:
        if not os.path.exists(f):
            continue
        size = os.path.getsize(f)
        t_ref = 1e30
        for _ in range(2):  # best of 2: the first call also pages the file in
            t0 = time.perf_counter()
            R.RefMatrix.read(f)
            t_ref = min(t_ref, time.perf_counter() - t0)
        best = 1e30
        for _ in range(5):
            t0 = time.perf_counter()
            info, _, _, _ = cb.mm_read_coo(f)
            best = min(best, time.perf_counter() - t0)
        print("| %s | %d | %d | %.3f | %.4f | %.0f | %.0fx |" % (os.path.basename(f), size, info["entries"], t_ref, best,
                                                                size / best / 1e6, t_ref / best))


if __name__ == "__main__":
    main()
