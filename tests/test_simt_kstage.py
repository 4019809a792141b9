"""The batched k-mer stage -- region_kmer_kernel + region_compact_kernel of breakmer_b200/csrc/region_kmers.cuh, the
product source -- on the host SIMT emulator (tests/sim/simt_host.h; 512 fibers per region, barriers at __syncthreads),
against the oracle's jellyfish-semantics counter and set algebra (oracle/kmers_py.py).  Shared-memory and global-memory
tables, empty / short / N-holding / lower-case records, regions without candidates, normal subtraction, several
interleavings of the threads.  Parity of the compiled kernel: tests/test_gpu_pipeline.py, tests/test_gpu_kmers.py."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import ROOT
from breakmer_b200 import synth
from oracle import kmers_py

SIM = os.path.join(ROOT, "tests", "sim")
BASES = "ACGT"


@pytest.fixture(scope="module")
def lib():
    import sim_util
    return ctypes.CDLL(sim_util.build_simt("libsimt_kstage.so", "simt_kstage.cpp"))


def _pack(per_region):
    """list (per region) of lists of sequences -> (bases, record offsets, region offsets)"""
    seqs = [s for reg in per_region for s in reg]
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    if seqs:
        np.cumsum([len(s) for s in seqs], out=off[1:])
    reg = np.zeros(len(per_region) + 1, dtype=np.int64)
    np.cumsum([len(r) for r in per_region], out=reg[1:])
    bases = np.frombuffer(("".join(seqs) + "\0").encode(), dtype=np.uint8).copy()
    return bases, off, reg


def run(lib, k, sc, reads, ref, normal=None, force_global=False, grid=3):
    R = len(sc)
    sets = [_pack(sc), _pack(reads), _pack([[r] for r in ref]), _pack(normal) if normal is not None else None]
    PB = ctypes.POINTER(ctypes.c_uint8) * 4
    PO = ctypes.POINTER(ctypes.c_int64) * 4
    pb, po, pr = PB(), PO(), PO()
    for i, s in enumerate(sets):
        if s is None:
            continue
        pb[i] = s[0].ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))
        po[i] = s[1].ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
        pr[i] = s[2].ctypes.data_as(ctypes.POINTER(ctypes.c_int64))
    n_sc = int(sets[0][1][-1])
    so_off = np.zeros(R + 1, dtype=np.int64)
    so_mer = np.zeros(n_sc + 1, dtype=np.uint64)
    so_cnt = np.zeros(n_sc + 1, dtype=np.uint32)
    rc = lib.simt_region_kmers(ctypes.c_int(R), ctypes.c_int(k), pb, po, pr, ctypes.c_int(1 if force_global else 0), ctypes.c_int(grid),
                               so_off.ctypes.data_as(ctypes.c_void_p), so_mer.ctypes.data_as(ctypes.c_void_p),
                               so_cnt.ctypes.data_as(ctypes.c_void_p))
    assert rc == 0
    out = []
    for r in range(R):
        a, b = int(so_off[r]), int(so_off[r + 1])
        mers = ["".join(BASES[(int(m) >> (2 * (k - 1 - i))) & 3] for i in range(k)) for m in so_mer[a:b]]
        assert mers == sorted(mers)                                 # ascending per region (A < C < G < T)
        out.append(dict(zip(mers, (int(c) for c in so_cnt[a:b]))))
    return out


def expected(k, sc, reads, ref, normal=None):
    return [kmers_py.sample_only(ref[r], reads[r], sc[r], k, normal[r] if normal is not None else None)[3] for r in range(len(sc))]


def _regions(wl, idx):
    regs = [synth.config_region(wl, i) for i in idx]
    sc = [[x[1] for x in r.sc_records] for r in regs]
    reads = [[x[1] for x in r.reads] for r in regs]
    ref = [r.ref_fwd for r in regs]
    normal = [[x[1] for x in r.normal_reads] for r in regs] if any(r.normal_reads for r in regs) else None
    return regs[0].k, sc, reads, ref, normal


@pytest.mark.parametrize("wl,idx", [("C2", (0, 1, 2, 76)), ("C3", (4, 5, 6)), ("C5", (0, 1, 2, 3, 50, 51)), ("C1", (0,))])
@pytest.mark.parametrize("force_global", [False, True])
def test_config_regions(lib, wl, idx, force_global):
    k, sc, reads, ref, normal = _regions(wl, idx)
    got = run(lib, k, sc, reads, ref, normal, force_global=force_global)
    exp = expected(k, sc, reads, ref, normal)
    assert got == exp
    assert sum(len(e) for e in exp) > 0


def test_ragged_and_odd_inputs(lib, monkeypatch):
    rng = random.Random(9)
    g = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    core = g(300)
    sc = [[core[10:90], "", "AC", core[100:160].lower(), core[150:200] + "N" + core[201:260]],   # empty, short, lower case, N
          [],                                                                                    # no candidates at all
          [g(40)],                                                                               # candidates, but no reads
          [core[0:50], core[0:50], "ACGTRYACGTACGTTTGACCA"],                                     # duplicates, IUPAC codes
          ["A" * 30 + core[5:40]]]                                                               # homopolymer run
    reads = [[core[0:100], core[50:150], core[50:150], "", core[120:220].lower(), core[180:280]],
             [g(100)],
             [],
             [core[0:60], "ACGTRYACGTACGTTTGACCA" * 2],
             ["A" * 60, "A" * 25 + core[5:60]]]
    ref = [core[0:40] + g(100), g(50), g(50), "", kmers_py.revcomp("A" * 30 + core[5:20]) if hasattr(kmers_py, "revcomp") else g(30)]
    normal = [[core[140:200]], [], [], [core[0:30]], []]
    for k in (3, 15, 21, 31):
        for order in (None, "reverse", "random:3"):
            if order:
                monkeypatch.setenv("SIMT_ORDER", order)
            else:
                monkeypatch.delenv("SIMT_ORDER", raising=False)
            assert run(lib, k, sc, reads, ref, normal, grid=2) == expected(k, sc, reads, ref, normal), (k, order)
    assert run(lib, 15, sc, reads, ref, None, force_global=True, grid=5) == expected(15, sc, reads, ref, None)


def test_windows_never_span_records_or_invalid_bases(lib):
    """adjacent pieces of one sequence as separate records: a window that (wrongly) spanned the cut would be a real k-mer
    of the other inputs, so it would show up in the result -- in every set, at every position of a thread's 4-window strip"""
    rng = random.Random(17)
    x = "".join(rng.choice("ACGT") for _ in range(400))
    y = "".join(rng.choice("ACGT") for _ in range(300))
    for k in (2, 3, 4, 5, 15, 22):
        for cut in range(100, 109):
            sc = [[x[40:cut], x[cut:260]],                       # candidates cut in two; the reads hold the uncut sequence
                  [x[30:300]],                                   # uncut candidates; the reads are cut
                  [x[30:300]],                                   # uncut candidates and reads; the reference is cut / holds an N
                  [x[40:cut] + "N" + x[cut + 1:260]]]
            reads = [[x[0:400], x[20:380]],
                     [x[0:cut], x[cut:400], x[10:cut], x[cut:390]],
                     [x[0:400], x[0:400]],
                     [x[0:400], x[0:400]]]
            ref = [y, y, x[0:cut] + "N" + x[cut + 1:150] + y, y]
            got = run(lib, k, sc, reads, ref, grid=4)
            assert got == expected(k, sc, reads, ref), (k, cut)
            normal = [[], [], [x[150:cut + 100], x[cut + 100:300]], []]
            assert run(lib, k, sc, reads, ref, normal, force_global=True) == expected(k, sc, reads, ref, normal), (k, cut)


def test_tile_boundaries(lib):
    """records and windows that straddle the 2,048-position tile of the window enumeration"""
    rng = random.Random(2)
    g = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    big = g(5000)
    sc = [[big[2030:2070], big[4090:4110], big[0:30]], [big[1000:3100]]]
    reads = [[big[0:2047], big[2047:2049], big[2049:4200], big[4080:4130]], [big[i:i + 101] for i in range(900, 3200, 7)]]
    ref = [g(2100), big[2000:2050] + g(2100)]
    for k in (4, 15, 31):
        assert run(lib, k, sc, reads, ref) == expected(k, sc, reads, ref), k
