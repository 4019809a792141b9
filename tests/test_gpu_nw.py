"""GPU parity: bk_nw_batch (the olc.nw drop-in) against the reference's own
outputs (golden) and against the oracle on fresh seeded pairs."""
import random

import numpy as np
import pytest

from conftest import golden
from oracle import nw_py

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from breakmer_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()


def _run(handle, pairs, want_aln=False):
    seqs = []
    pa, pb = [], []
    for a, b in pairs:
        pa.append(len(seqs)); seqs.append(a)
        pb.append(len(seqs)); seqs.append(b)
    return handle.nw_batch(seqs, pa, pb, want_aln=want_aln)


def test_golden_pairs_both_directions(handle):
    cases = golden("nw_golden.json")["cases"]
    out, _ = _run(handle, [(c["seq1"], c["seq2"]) for c in cases])
    for c, o in zip(cases, out):
        assert list(o[:5]) == c["out"][2:], (c["seq1"], c["seq2"])
        assert list(o[5:]) == list(nw_py.nw_fast(c["seq2"], c["seq1"])[2:])


def test_golden_alignment_strings(handle):
    cases = golden("nw_golden.json")["cases"]
    out, alns = _run(handle, [(c["seq1"], c["seq2"]) for c in cases], want_aln=True)
    for c, o, (a1, a2) in zip(cases, out, alns):
        assert [a1, a2] + list(o[:5]) == c["out"]


def test_random_pairs_incl_long_and_n(handle):
    rng = random.Random(11)
    pairs = []
    for t in range(600):
        la = rng.choice([1, 2, 17, 100, 101, 128, 129, 190, 256, 257, 300, 513, 700])
        lb = rng.choice([1, 3, 31, 32, 33, 64, 100, 150, 260, 400])
        g = "".join(rng.choice("ACGT") for _ in range(la + lb))
        a = g[:la]
        if t % 3 == 0:
            b = "".join(rng.choice("ACGTN") for _ in range(lb))
        elif t % 3 == 1:
            ov = rng.randint(1, min(la, lb))
            b = (a[la - ov:] + g[la:])[:lb]
        else:
            b = g[max(0, la - lb // 2):][:lb]
        b = "".join(c if rng.random() > 0.02 else rng.choice("ACGTN") for c in b)
        pairs.append((a, b) if t % 2 else (b, a))
    out, _ = _run(handle, pairs)
    for (a, b), o in zip(pairs, out):
        assert list(o[:5]) == list(nw_py.nw_fast(a, b)[2:]), (a, b)
        assert list(o[5:]) == list(nw_py.nw_fast(b, a)[2:]), (a, b)


def test_max_packed_length_and_empty_sequence(handle):
    from breakmer_b200 import _lib
    rng = random.Random(3)
    a = "".join(rng.choice("ACGT") for _ in range(4095))
    b = a[3000:] + "".join(rng.choice("ACGT") for _ in range(500))
    out, _ = _run(handle, [(a, b), (b, a)])
    assert list(out[0][:5]) == list(nw_py.nw_fast(a, b)[2:])
    assert list(out[1][:5]) == list(nw_py.nw_fast(b, a)[2:])
    with pytest.raises(_lib.BreakmerError) as e:
        _run(handle, [("", "ACGT")])
    assert e.value.code == _lib.BK_ERR_EMPTY_SEQ      # the reference raises NameError here


def test_score_pass_traceback_equals_packed_cell_kernel(handle, monkeypatch):
    """Both DP kernels of nw.cuh on the same pairs: the score pass + warp-parallel traceback (default where its table
    fits) and the packed-cell kernel that carries origins forward (BK_NW_PACKED=1 forces it everywhere)."""
    rng = random.Random(29)
    pairs = []
    for t in range(1500):
        la = rng.randint(1, 128)                       # columns: the trace path takes one column block of 4 per lane
        lb = rng.choice([1, 2, 5, 40, 99, 100, 101, 160, 230, 400, 800, 1216, 1217, 1300])
        g = "".join(rng.choice("ACGT") for _ in range(la + lb + 50))
        mode = t % 5
        if mode == 0:                                  # suffix of b overlaps prefix of a
            ov = rng.randint(1, min(la, lb))
            b = (g[la:la + lb - ov] + g[:ov])[:lb]
            a = g[:la]
        elif mode == 1:                                # contained
            a = g[10:10 + la]; b = g[:lb] if lb > la + 10 else g[:la + 20]
        elif mode == 2:                                # low-complexity / repeats: many ties
            unit = "".join(rng.choice("AC") for _ in range(rng.randint(1, 4)))
            a = (unit * 200)[:la]; b = (unit * 700)[:lb]
        elif mode == 3:                                # unrelated
            a = g[:la]; b = "".join(rng.choice("ACGTN") for _ in range(lb))
        else:                                          # overlap with indels
            a = g[:la]
            b = list(g[max(0, la - 60):][:lb])
            for _ in range(rng.randint(0, 4)):
                if b:
                    p = rng.randrange(len(b))
                    if rng.random() < 0.5:
                        del b[p]
                    else:
                        b.insert(p, rng.choice("ACGT"))
            b = "".join(b) or "A"
        if rng.random() < 0.3:
            b = "".join(c if rng.random() > 0.03 else rng.choice("ACGTN") for c in b)
        pairs.append((a, b))
    fast, _ = _run(handle, pairs)
    monkeypatch.setenv("BK_NW_PACKED", "1")
    packed, _ = _run(handle, pairs)
    monkeypatch.delenv("BK_NW_PACKED")
    assert np.array_equal(fast, packed)
    for (a, b), o in list(zip(pairs, fast))[::7]:
        assert list(o[:5]) == list(nw_py.nw_fast(a, b)[2:]), (a, b)
        assert list(o[5:]) == list(nw_py.nw_fast(b, a)[2:]), (a, b)


def _long_pairs():
    """pairs with a sequence above the packed-cell kernels' 4095 bases (olc.nw has no limit, olc.py:40-52): overlaps,
    containment, unrelated, one base against a long one, repeats (ties), both long"""
    rng = random.Random(41)
    g = "".join(rng.choice("ACGT") for _ in range(12000))
    mut = lambda s, p: "".join(c if rng.random() > p else rng.choice("ACGTN") for c in s)
    rep = ("AC" * 3000)
    return [
        (g[:4096], g[3500:4700]),                 # suffix of seq1 overlaps the prefix of seq2, seq1 one base over the limit
        (g[3500:4700], g[:4096]),
        (g[:5000], mut(g[4200:5600], 0.03)),      # overlap with mismatches
        (g[:4600], g[1000:1300]),                 # seq2 contained in seq1
        (g[1000:1300], g[:4600]),
        (g[:4100], g[6000:6300]),                 # unrelated
        ("A", g[:4200]),
        (g[:4200], "T"),
        (rep[:4300], rep[1:901]),                 # low complexity: ties everywhere
        (g[:4500], mut(g[300:4900], 0.01)),       # both above the limit
        (g[:100], g[50:150]),                     # an ordinary pair in the same call
    ]


def test_pairs_above_4095_bases(handle):
    pairs = _long_pairs()
    out, _ = _run(handle, pairs)
    for (a, b), o in zip(pairs, out):
        assert list(o[:5]) == list(nw_py.nw_fast(a, b)[2:]), (len(a), len(b))
        assert list(o[5:]) == list(nw_py.nw_fast(b, a)[2:]), (len(a), len(b))


def test_alignment_strings_above_4095_bases(handle):
    pairs = _long_pairs()[:4] + _long_pairs()[6:9] + _long_pairs()[10:]
    out, alns = _run(handle, pairs, want_aln=True)
    for (a, b), o, (a1, a2) in zip(pairs, out, alns):
        assert (a1, a2) + tuple(o[:5]) == tuple(nw_py.nw_fast(a, b)), (len(a), len(b))
        assert list(o[5:]) == list(nw_py.nw_fast(b, a)[2:]), (len(a), len(b))
