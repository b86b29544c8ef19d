"""csrc/myers.cuh compiled for the host: the bit-parallel distance equals the DP oracle for every
length pair the path accepts (1..32 in one word, 1..64 in two) -- the device runs the same code."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import orc
from pacybara_b200 import hostdata

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_emul")


@pytest.fixture(scope="module")
def lib():
    so = os.path.join(HERE, "libmyers_emul.so")
    src = os.path.join(HERE, "myers_emul.cpp")
    inc = os.path.join(HERE, "..", "..", "pacybara_b200", "csrc")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(inc, "myers.cuh"))):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O1", "-shared", "-fPIC", "-x", "c++", "-I", inc, "-o", so, src], check=True)
    return C.CDLL(so)


def _mutate(rng, s, n_edits):
    s = list(s)
    for _ in range(n_edits):
        op = rng.integers(0, 3)
        if op == 0 and s:
            s[rng.integers(0, len(s))] = "ACGT"[rng.integers(0, 4)]
        elif op == 1:
            s.insert(rng.integers(0, len(s) + 1), "ACGT"[rng.integers(0, 4)])
        elif s and len(s) > 1:
            del s[rng.integers(0, len(s))]
    return "".join(s)


@pytest.mark.parametrize("maxlen,use64", [(32, 0), (64, 1), (32, 1)])
def test_myers_equals_dp(lib, maxlen, use64):
    rng = np.random.default_rng(5 + maxlen + use64)
    seqs = []
    for _ in range(1500):
        n = int(rng.integers(1, maxlen + 1))
        a = "".join("ACGT"[x] for x in rng.integers(0, 4, size=n))
        b = _mutate(rng, a, int(rng.integers(0, 8)))[:maxlen]
        seqs += [a.encode(), (b or "A").encode()]
    # extremes: full-length words, single bases, homopolymers
    seqs += [b"A" * maxlen, b"T" * maxlen, b"A", b"T", b"ACGT" * (maxlen // 4), b"TGCA" * (maxlen // 4)]
    mat, lens = hostdata.strings_to_matrix(seqs, 64)
    n = len(seqs)
    pi = rng.integers(0, n, size=6000).astype(np.int32)
    pj = rng.integers(0, n, size=6000).astype(np.int32)
    pi[:n // 2] = np.arange(0, n - 1, 2)[: n // 2]         # the related pairs
    pj[:n // 2] = np.arange(1, n, 2)[: n // 2]
    out = np.zeros(pi.shape[0], dtype=np.int32)
    lib.emul_myers(mat.ctypes.data_as(C.c_void_p), lens.ctypes.data_as(C.c_void_p), C.c_int32(64),
                   C.c_int64(pi.shape[0]), pi.ctypes.data_as(C.c_void_p), pj.ctypes.data_as(C.c_void_p),
                   C.c_int32(use64), out.ctypes.data_as(C.c_void_p))
    want = np.array([orc.lev(seqs[a], seqs[b]) for a, b in zip(pi, pj)], dtype=np.int32)
    assert np.array_equal(out, want)
