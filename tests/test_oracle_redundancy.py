"""Read-redundancy row on the CPU: the oracle restatement against the golden
vectors generated from the reference's own same_reads / subseq / sim_seqs / read_batch.check_mer_read
(oracle/make_golden_redundancy.py), and the host replay of bk_dedup_reads (csrc/dedup.cuh, compiled
for the host with the score table supplied by the oracle's nw) against the same vectors."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden
from oracle import nw_py, redundancy_py as R

G = golden("redundancy_cases.json")


def test_pairs_against_reference():
    for c in G["pairs"]:
        a, b = c["seq1"], c["seq2"]
        assert R.same_reads(a, b) == c["same_reads"] == c["same_reads_mm2"]
        assert list(R.subseq(a, b, R.SUBSEQ_FRAC_LIVE)) == c["subseq_live"]
        assert list(R.subseq(a, b, R.SUBSEQ_FRAC_MM2)) == c["subseq_mm2"]
        got = []
        for red in (False, True):
            br = R.b_read("x", b)
            br.redundant = red
            got.append(R.sim_seqs(a, br))
        assert got == c["sim_seqs"]


def test_sim_seqs_is_true_for_unrelated_reads():
    # the tuple-truthiness quirk (R1): pinned by the golden file, stated here explicitly
    assert any(c["sim_seqs"][0] and not c["same_reads"] and not c["subseq_mm2"][0] for c in G["pairs"])
    assert all(c["sim_seqs"] == [True, False] for c in G["pairs"])


def test_pure_python_nw_gives_the_same_predicates():
    for c in G["pairs"][:12]:
        assert list(R.subseq(c["seq1"], c["seq2"], nw=nw_py.nw)) == c["subseq_mm2"]
        assert R.same_reads(c["seq1"], c["seq2"], nw=nw_py.nw) == c["same_reads"]


def test_batches_against_reference():
    for b in G["batches"]:
        checks, kept, redundant, deleted = R.dedup_batch([tuple(r) for r in b["reads"]])
        assert checks == b["checks"]
        assert kept == b["kept"]
        assert redundant == b["redundant"]
        assert deleted == b["deleted"]


def test_empty_sequence_raises_like_the_reference():
    with pytest.raises(NameError):
        R.dedup_batch([("a", "ACGT", 0), ("b", "", 1)])


# ---- the product's host replay (csrc/dedup.cuh), with the oracle's nw standing in for the kernel ----------

@pytest.fixture(scope="module")
def sim():
    so = os.path.join(ROOT, "tests", "sim", "libdedup_sim.so")
    src = os.path.join(ROOT, "tests", "sim", "dedup_sim.cpp")
    deps = [src, os.path.join(ROOT, "breakmer_b200", "csrc", "dedup.cuh"), os.path.join(ROOT, "include", "breakmer_b200.h")]
    if not os.path.isfile(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", so, src])
    return ctypes.CDLL(so)


ALIGN_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32), ctypes.c_int64,
                            ctypes.POINTER(ctypes.c_int32))


def replay_many(sim, batches, frac):
    """batches: [[(id, seq, pos)]] -> ([(checks, kept, redundant, deleted)], alignments, rounds) through dedup_run,
    with the oracle's nw standing in for the kernel."""
    reads = [r for b in batches for r in b]
    n = len(reads)
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum([len(r[1]) for r in reads], out=off[1:])
    pos = np.array([r[2] for r in reads], dtype=np.int32)
    boff = np.zeros(len(batches) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in batches], out=boff[1:])
    asked = []

    def align(pa, pb, cnt, out):
        for p in range(cnt):
            i, j = pa[p], pb[p]
            assert i < j
            asked.append((i, j))
            f = nw_py.nw_fast(reads[i][1], reads[j][1])[2:] + nw_py.nw_fast(reads[j][1], reads[i][1])[2:]
            for t in range(10):
                out[p * 10 + t] = f[t]
        return 0

    check = np.full(n, 7, dtype=np.uint8)
    flags = np.full(n, 255, dtype=np.uint8)
    n_pairs, n_rounds = ctypes.c_int64(), ctypes.c_int()
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    sim.dedup_sim_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_double,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int64),
                                  ctypes.POINTER(ctypes.c_int), ALIGN_FN]
    rc = sim.dedup_sim_run(p(off), p(pos), p(boff), len(batches), frac, p(check), p(flags), ctypes.byref(n_pairs),
                           ctypes.byref(n_rounds), ALIGN_FN(align))
    assert rc == 0
    assert n_pairs.value == len(asked) == len(set(asked))            # no alignment is computed twice
    out = []
    for bi, b in enumerate(batches):
        lo = int(boff[bi])
        ids = [r[0] for r in b]
        fl = flags[lo:lo + len(b)]
        out.append(([bool(c) for c in check[lo:lo + len(b)]], [ids[i] for i in range(len(b)) if fl[i] & 1],
                    [ids[i] for i in range(len(b)) if fl[i] & 2], sorted(ids[i] for i in range(len(b)) if fl[i] & 4)))
    return out, n_pairs.value, n_rounds.value


def replay(sim, reads, frac):
    return replay_many(sim, [reads], frac)[0][0]


def test_host_replay_against_reference(sim):
    for b in G["batches"]:
        checks, kept, redundant, deleted = replay(sim, [tuple(r) for r in b["reads"]], R.SUBSEQ_FRAC_MM2)
        assert (checks, kept, redundant, deleted) == (b["checks"], b["kept"], b["redundant"], b["deleted"])


def random_batch(rng, n, tag):
    g = "".join(rng.choice("ACGT") for _ in range(400))
    reads = []
    for i in range(n):
        rl = rng.choice([100, 100, 80, 60, 45])
        pos = rng.randint(0, rl - 15)
        seq = list(g[200 - pos:200 - pos + rl])
        for _ in range(rng.choice([0, 0, 1, 2, 5])):
            x = rng.randrange(len(seq))
            seq[x] = rng.choice("ACGTN")
        if reads and rng.random() < 0.2:
            prev = reads[-1]
            cut = rng.randint(0, 6)
            seq, pos = list(prev[1][cut:len(prev[1]) - rng.randint(0, 6)]), prev[2] + 100 + i    # contained, fresh position
        reads.append(("%s_%d" % (tag, i), "".join(seq), pos))
    return reads


def test_host_replay_against_oracle_on_random_batches(sim):
    rng = random.Random(77)
    seen = set()
    for t in range(40):
        reads = random_batch(rng, rng.randint(1, 25), "t%d" % t)
        for frac in (R.SUBSEQ_FRAC_MM2, R.SUBSEQ_FRAC_LIVE):
            exp = R.dedup_batch(reads, frac)
            got = replay(sim, reads, frac)
            assert got == (exp[0], exp[1], exp[2], exp[3])
            seen.add((bool(exp[2]), len(exp[3]) > 0))
    assert (True, True) in seen          # the redundant-flag arm was exercised


def test_host_replay_many_batches_in_shared_rounds(sim):
    """All batches of a call share the launches: the golden batches and random ones in one run, fewer alignments
    than all ordered pairs, and only a handful of rounds."""
    rng = random.Random(5)
    batches = [[tuple(r) for r in b["reads"]] for b in G["batches"]] + [random_batch(rng, 40, "m%d" % t) for t in range(12)]
    got, n_pairs, n_rounds = replay_many(sim, batches, R.SUBSEQ_FRAC_MM2)
    for b, g in zip(batches, got):
        exp = R.dedup_batch(b, R.SUBSEQ_FRAC_MM2)
        assert g == (exp[0], exp[1], exp[2], exp[3])
    all_pairs = sum(len(b) * (len(b) - 1) // 2 for b in batches)
    assert n_pairs < all_pairs // 3
    assert 1 <= n_rounds <= 8


def test_host_replay_long_run_of_dropped_reads(sim):
    """More reads dropped in a row than the band and the window hold: the chain asks again, round after round."""
    rng = random.Random(6)
    s = "".join(rng.choice("ACGT") for _ in range(100))
    inner = [("in%d" % i, s[1 + i % 3:90 - i % 5], 100 + i) for i in range(30)]       # each inside the opener
    reads = [("open", s, 0)] + inner + [("tail", "".join(rng.choice("ACGT") for _ in range(80)), 500)]
    exp = R.dedup_batch(reads, R.SUBSEQ_FRAC_MM2)
    assert exp[1] == ["open", "tail"]
    got, n_pairs, n_rounds = replay_many(sim, [reads], R.SUBSEQ_FRAC_MM2)
    assert got[0] == (exp[0], exp[1], exp[2], exp[3])
    assert n_rounds >= 4


def test_invariants_of_a_dedup_result(sim):
    """Size-independent properties of check_mer_read's bookkeeping, on batches larger than the oracle comparison uses:
    every read is either appended to the batch or in the delete set (a kept read flagged redundant is in both), the
    opener is always kept, check is true exactly for the appended reads, and two kept, unflagged reads never share a
    seed position (the sim_seqs quirk drops the second one)."""
    rng = random.Random(11)
    batches = [random_batch(rng, rng.randint(50, 200), "p%d" % t) for t in range(10)]
    got, n_pairs, n_rounds = replay_many(sim, batches, R.SUBSEQ_FRAC_MM2)
    for b, (checks, kept, redundant, deleted) in zip(batches, got):
        ids = [r[0] for r in b]
        pos = {r[0]: r[2] for r in b}
        assert kept[0] == ids[0] and checks[0]
        assert [i for i, c in zip(ids, checks) if c] == kept
        assert set(kept) | set(deleted) == set(ids)
        assert set(kept) & set(deleted) == set(redundant)
        live = [pos[i] for i in kept if i not in set(redundant)]
        assert len(live) == len(set(live))
    assert n_pairs < sum(len(b) * (len(b) - 1) // 2 for b in batches) // 5
