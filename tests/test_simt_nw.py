"""The warp DP kernels of breakmer_b200/csrc/nw.cuh -- the product source, unmodified -- run on a 32-lane host emulator
(tests/sim/simt_host.h: one fiber per lane, barriers at the warp collectives) and compared with the reference's own
`olc.nw` outputs (tests/golden/nw_golden.json) and with the oracle on seeded pairs.  This is a CPU-side check of the
kernel SOURCE (index arithmetic, table layout, traceback, lane exchange, missing __syncwarp()); parity of the compiled
sm_100a kernels is tests/test_gpu_nw.py."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import nw_py

SIM = os.path.join(ROOT, "tests", "sim")


@pytest.fixture(scope="module")
def lib():
    import sim_util
    return ctypes.CDLL(sim_util.build_simt("libsimt_nw.so", "simt_nw.cpp"))


def run(lib, cs, rs, mode):
    out = (ctypes.c_int * 10)()
    rc = lib.simt_nw_dual(cs.encode(), len(cs), rs.encode(), len(rs), mode, out)
    assert rc == 0, "simt_nw_dual rc %d (lanes disagree if < -1)" % rc
    return list(out)


def expected(cs, rs):
    return list(nw_py.nw_fast(cs, rs)[2:]) + list(nw_py.nw_fast(rs, cs)[2:])


def lazy_reads(exp, m, n):
    """which traceback origins contig.check_align reads (sv_assembly.py:449-504), as nw.cuh's LAZY states them"""
    sa, sb = exp[4], exp[9]
    mn = min(m, n)
    low_a, low_b = 4 * sa < mn, 4 * sb < mn
    if sa == sb:
        return (not low_a, not low_a)
    first_a = sa > sb
    if low_a if first_a else low_b:
        return (False, False)
    span = (m - exp[1]) if first_a else (n - exp[6])
    bad1 = 200 * (sa if first_a else sb) < 179 * span
    other = bad1 and not (low_b if first_a else low_a)
    return (True, other) if first_a else (other, True)


def check(lib, cs, rs, modes=(0, 1, 2)):
    exp = expected(cs, rs)
    for mode in modes:
        got = run(lib, cs, rs, mode)
        if mode == 1:
            ra, rb = lazy_reads(exp, len(cs), len(rs))
            keep = [0, 2, 4, 5, 7, 9] + ([1, 3] if ra else []) + ([6, 8] if rb else [])
            assert [got[i] for i in keep] == [exp[i] for i in keep], (mode, cs, rs)
        else:
            assert got == exp, (mode, cs, rs)


def test_reference_golden_pairs(lib):
    """every pair the reference itself aligned for the golden file, both argument orders, all kernels"""
    cases = golden("nw_golden.json")["cases"]
    n = 0
    for c in cases:
        a, b = c["seq1"], c["seq2"]
        if not a or not b:
            continue
        assert expected(a, b)[:5] == c["out"][2:]             # the oracle agrees with the reference (pinned elsewhere too)
        check(lib, a, b)
        if n % 4 == 0:
            check(lib, b, a, modes=(0, 3))
        n += 1
    assert n > 350


def test_seeded_pairs_all_kernels(lib):
    """read x contig shapes of the assembler (overhangs at either end, repeats, indels, N), column counts on both sides
    of every lane / block boundary"""
    rng = random.Random(23)
    pairs = []
    for t in range(260):
        la = rng.choice([1, 2, 3, 4, 5, 31, 32, 33, 63, 64, 65, 97, 100, 101, 124, 125, 126, 127, 128, 129, 150, 257, 300])
        lb = rng.choice([1, 2, 3, 17, 64, 100, 150, 199, 200, 201, 260, 400])
        g = "".join(rng.choice("ACGT") for _ in range(la + lb))
        a = g[:la]
        kind = t % 5
        if kind == 0:
            b = "".join(rng.choice("ACGTN") for _ in range(lb))
        elif kind == 1:
            ov = rng.randint(1, min(la, lb))
            b = (a[la - ov:] + g[la:])[:lb]                    # b starts inside a and runs past its end
        elif kind == 2:
            ov = rng.randint(1, min(la, lb))
            b = (g[la:la + lb - ov] + a[:ov])                   # b ends inside a's start
        elif kind == 3:
            unit = "".join(rng.choice("ACGT") for _ in range(rng.choice([1, 2, 3, 7])))
            a = (unit * (la // len(unit) + 1))[:la]
            b = (unit * (lb // len(unit) + 1))[:lb]            # repeats: ties everywhere
        else:
            b = g[max(0, la - lb // 2):][:lb]
            cut = rng.randint(0, max(0, len(b) - 4))
            b = b[:cut] + b[cut + rng.randint(1, 3):] if rng.random() < 0.5 else b[:cut] + "GG" + b[cut:]
        b = "".join(c if rng.random() > 0.02 else rng.choice("ACGTN") for c in b) or "A"
        pairs.append((a, b))
    for a, b in pairs:
        check(lib, a, b)


def test_other_interleavings(lib, monkeypatch):
    rng = random.Random(41)
    for order in ("reverse", "random:5"):
        monkeypatch.setenv("SIMT_ORDER", order)
        for _ in range(30):
            la, lb = rng.choice([37, 100, 101, 128]), rng.choice([60, 150, 260])
            g = "".join(rng.choice("ACGT") for _ in range(la + lb))
            ov = rng.randint(5, min(la, lb))
            check(lib, g[:la], (g[la - ov:la] + g[la:])[:lb])


def test_long_sequences_take_the_packed_kernel(lib):
    rng = random.Random(5)
    a = "".join(rng.choice("ACGT") for _ in range(700))
    b = a[500:] + "".join(rng.choice("ACGT") for _ in range(300))
    assert not lib.simt_nw_trace_fits(len(a), len(b)) and lib.simt_nw_trace_fits(100, 200)
    check(lib, a, b, modes=(0, 1, 3))
    check(lib, b, a, modes=(0, 2))
    c = "".join(rng.choice("ACGT") for _ in range(100))
    d = "".join(rng.choice("ACGT") for _ in range(1300))      # 100 columns, but more rows than the score table holds
    assert not lib.simt_nw_trace_fits(len(c), len(d))
    check(lib, c, d[:600] + c[40:] + d[600:], modes=(0, 1))


# ---- nw_batch_kernel: the device side of bk_nw_batch (four warps per block, pairs strided over the warps) ------------
@pytest.fixture(scope="module")
def batch_lib():
    import sim_util
    return ctypes.CDLL(sim_util.build_simt("libsimt_nw_batch.so", "simt_nw_batch.cpp"))


def run_batch(lib, pairs, want_aln, grid=2, use_tab=True, long_kernel=False):
    seqs, pa, pb = [], [], []
    for a, b in pairs:
        pa.append(len(seqs)); seqs.append(a)
        pb.append(len(seqs)); seqs.append(b)
    off = np.zeros(len(seqs) + 1, dtype=np.int64)
    np.cumsum([len(s) for s in seqs], out=off[1:])
    blob = np.frombuffer(("".join(seqs) + "\0").encode(), dtype=np.uint8).copy()
    pa = np.array(pa, dtype=np.int32); pb = np.array(pb, dtype=np.int32)
    n = len(pairs)
    out = np.zeros(n * 10, dtype=np.int32)
    aln_off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(a) + len(b) for a, b in pairs], out=aln_off[1:])
    a1 = np.zeros(int(aln_off[-1]) + 1, dtype=np.uint8); a2 = np.zeros_like(a1)
    alen = np.zeros(n, dtype=np.int32)
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
    if long_kernel:
        rc = lib.simt_nw_long(p(blob), p(off), ctypes.c_int(len(seqs)), p(pa), p(pb), ctypes.c_int64(n), p(out),
                              ctypes.c_int(1 if want_aln else 0), p(a1), p(a2), p(aln_off), p(alen))
    else:
        rc = lib.simt_nw_batch(p(blob), p(off), ctypes.c_int(len(seqs)), p(pa), p(pb), ctypes.c_int64(n), ctypes.c_int(grid),
                               ctypes.c_int(1 if use_tab else 0), p(out), ctypes.c_int(1 if want_aln else 0), p(a1), p(a2), p(aln_off), p(alen))
    assert rc == 0
    res = []
    for i in range(n):
        s = int(aln_off[i])
        res.append((bytes(a1[s:s + alen[i]]).decode(), bytes(a2[s:s + alen[i]]).decode(), [int(v) for v in out[10 * i:10 * i + 10]]))
    return res


def test_batch_kernel_alignment_strings_of_the_reference(batch_lib):
    """the full tuple olc.nw returns -- both alignment strings and the five integers -- for every golden pair"""
    cases = [c for c in golden("nw_golden.json")["cases"] if c["seq1"] and c["seq2"]]
    got = run_batch(batch_lib, [(c["seq1"], c["seq2"]) for c in cases], want_aln=True, grid=3)
    for c, (a1, a2, o) in zip(cases, got):
        assert [a1, a2] + o[:5] == c["out"], (c["seq1"], c["seq2"])
        assert o[5:] == list(nw_py.nw_fast(c["seq2"], c["seq1"])[2:])


def test_batch_kernel_without_strings(batch_lib):
    cases = [c for c in golden("nw_golden.json")["cases"] if c["seq1"] and c["seq2"]][::3]
    for use_tab in (True, False):
        got = run_batch(batch_lib, [(c["seq1"], c["seq2"]) for c in cases], want_aln=False, grid=1, use_tab=use_tab)
        for c, (_a1, _a2, o) in zip(cases, got):
            assert o[:5] == c["out"][2:]


# ---- nw_long_kernel: bk_nw_batch's kernel for pairs above 4095 bases (one block per direction, 32-bit anti-diagonals) ---
def test_long_kernel_on_the_reference_golden_pairs(batch_lib):
    """the kernel has no LOWER length bound, so every golden pair of the reference runs through it: full tuple of
    nw(seq1, seq2) incl. both alignment strings, and the five integers of nw(seq2, seq1)"""
    cases = [c for c in golden("nw_golden.json")["cases"] if c["seq1"] and c["seq2"]]
    got = run_batch(batch_lib, [(c["seq1"], c["seq2"]) for c in cases], want_aln=True, long_kernel=True)
    for c, (a1, a2, o) in zip(cases, got):
        assert [a1, a2] + o[:5] == c["out"], (c["seq1"], c["seq2"])
        assert o[5:] == list(nw_py.nw_fast(c["seq2"], c["seq1"])[2:])


def test_long_kernel_equals_the_warp_kernels_on_seeded_pairs(batch_lib):
    """no pointer table (the stop position is carried forward with the score): ties, repeats, N, one-base sequences,
    lengths around the 256-thread stride"""
    rng = random.Random(17)
    pairs = []
    for t in range(160):
        la = rng.choice([1, 2, 3, 31, 100, 255, 256, 257, 300, 513, 700])
        lb = rng.choice([1, 2, 33, 100, 150, 256, 258, 400])
        g = "".join(rng.choice("ACGT") for _ in range(la + lb))
        a = g[:la]
        mode = t % 4
        if mode == 0:
            b = "".join(rng.choice("ACGTN") for _ in range(lb))
        elif mode == 1:
            ov = rng.randint(1, min(la, lb))
            b = (a[la - ov:] + g[la:])[:lb]
        elif mode == 2:
            unit = "".join(rng.choice("AC") for _ in range(rng.randint(1, 3)))
            a = (unit * 800)[:la]; b = (unit * 800)[1:1 + lb]
        else:
            b = g[max(0, la - lb // 2):][:lb]
        b = "".join(c if rng.random() > 0.02 else rng.choice("ACGTN") for c in b)
        pairs.append((a, b) if t % 2 else (b, a))
    got = run_batch(batch_lib, pairs, want_aln=False, long_kernel=True)
    for (a, b), (_a1, _a2, o) in zip(pairs, got):
        assert o[:5] == list(nw_py.nw_fast(a, b)[2:]), (a, b)
        assert o[5:] == list(nw_py.nw_fast(b, a)[2:]), (a, b)


def test_long_kernel_on_tie_heavy_short_pairs(batch_lib):
    """small alphabets and short sequences: nearly every cell is a tie, so the pointer priority diag > up > left and the
    stop position carried forward with the score are exercised at every branch; full tuple incl. strings (3,000 such pairs
    were run once by hand: 0 mismatches)"""
    rng = random.Random(2026)
    pairs = []
    for _ in range(800):
        alpha = rng.choice(["A", "AC", "ACG", "ACGT", "ACGTN"])
        pairs.append(("".join(rng.choice(alpha) for _ in range(rng.randint(1, 40))),
                      "".join(rng.choice(alpha) for _ in range(rng.randint(1, 40)))))
    got = run_batch(batch_lib, pairs, want_aln=True, long_kernel=True)
    for (a, b), (a1, a2, o) in zip(pairs, got):
        assert (a1, a2) + tuple(o[:5]) == tuple(nw_py.nw_fast(a, b)), (a, b)
        assert o[5:] == list(nw_py.nw_fast(b, a)[2:]), (a, b)
