"""CPU suite: the C-ABI library loads, exports every symbol include/cask_b200.h declares, and its
host-only entry points behave (no compute call is made without a GPU)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import cask_b200
    from cask_b200 import build
    build.build()
    return cask_b200


def test_every_declared_symbol_is_exported(lib):
    header = open(os.path.join(ROOT, "include", "cask_b200.h")).read()
    declared = set(re.findall(r"\b(cask_b200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 28
    L = C.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "missing export " + name
    assert declared == set(lib.EXPORTED_SYMBOLS)


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "cask_b200.h"\nint main(void){cask_b200_design d; (void)d; return sizeof(cask_b200_partition_info) != 56;}\n')
    exe = tmp_path / "t"
    assert os.system("gcc -std=c99 -Wall -Werror -I%s/include %s -o %s" % (ROOT, src, exe)) == 0
    assert os.system(str(exe)) == 0


def test_shard_rows_follow_reference_striping(lib):
    """rows_per = n / world, remainder to the last (Spmv.cpp:334,353-364)."""
    for n, w in ((16777216, 8), (17, 4), (5, 8), (100, 3)):
        tot = 0
        for r in range(w):
            r0, nr = lib.shard_rows(n, w, r)
            assert r0 == (n // w) * r
            assert nr == (n // w if r < w - 1 else n - (n // w) * (w - 1))
            tot += nr
        assert tot == n
    with pytest.raises(lib.CaskError):
        lib.shard_rows(10, 2, 2)


def test_synth_counts(lib):
    assert lib.synth_rows(lib.SYNTH_POISSON2D, 4096) == 16777216
    assert lib.synth_nnz(lib.SYNTH_POISSON2D, 4096, 0, 16777216) == 83869696          # C2
    assert lib.synth_nnz(lib.SYNTH_POISSON3D27, 256, 0, 256 ** 3) == 766 ** 3         # C4: 449 455 096
    assert lib.synth_nnz(lib.SYNTH_CONVDIFF3D7, 512, 0, 512 ** 3) == 937951232        # C5
    from oracle import oraclebind as O
    for kind, gen, N in ((0, O.gen_poisson2d, 13), (1, O.gen_poisson3d27, 7), (2, O.gen_convdiff3d7, 6)):
        n, rp, ci, va = gen(N)
        for r0, nr in ((0, n), (5, 17), (n - 3, 3), (n // 2, 0)):
            assert lib.synth_nnz(kind, N, r0, nr) == rp[r0 + nr] - rp[r0]


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.CaskError) as e:
        lib.Context(0)
    assert e.value.code == lib.ERR_NO_DEVICE and "no CPU fallback" in e.value.message


def test_product_never_touches_the_oracle():
    """The product path must not import, link or dlopen anything under oracle/."""
    pkg = os.path.join(ROOT, "cask_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".inl")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                for pat in (r"#\s*include[^\n]*oracle", r"^\s*(from|import)\s+oracle", r"import[^\n]*oraclebind",
                            r"dlopen[^\n]*oracle", r"libcask_oracle", r"caskref", r"-l\s*cask_oracle",
                            # nor the host emulation of the device-logic backend (tests/emu): test infrastructure only
                            r"#\s*include[^\n]*dev_host", r"#\s*include[^\n]*tests/", r"libcask_emu"):
                    assert not re.search(pat, text, re.M), (os.path.join(dirpath, f), pat)
    out = os.popen("ldd %s" % os.path.join(pkg, "libcask_b200.so")).read()
    assert "oracle" not in out and "caskref" not in out and "emu" not in out
