"""GPU parity of the read-redundancy row: bk_dedup_reads and the predicates
over bk_nw_batch against the reference's own outputs (tests/golden/redundancy_cases.json) and against the
oracle on fresh seeded batches."""
import random

import pytest

from conftest import golden
from oracle import redundancy_py as R
from test_oracle_redundancy import random_batch

pytestmark = pytest.mark.gpu

G = golden("redundancy_cases.json")


class Read:
    def __init__(self, rid, seq):
        self.id = rid
        self.seq = seq


def as_batch(reads):
    return [(Read(r[0], r[1]), r[2]) for r in reads]


def unpack(res):
    return (res["checks"], [r.id for r in res["kept"]], [r.id for r in res["redundant"]], sorted(res["deleted"]))


def test_predicates_against_reference():
    from breakmer_b200 import sv_assembly_mm2 as mm2
    pairs = [(c["seq1"], c["seq2"]) for c in G["pairs"]]
    assert mm2.same_reads_batch(pairs) == [c["same_reads"] for c in G["pairs"]]
    assert [list(t) for t in mm2.subseq_batch(pairs)] == [c["subseq_mm2"] for c in G["pairs"]]
    assert [list(t) for t in mm2.subseq_batch(pairs, frac=0.85)] == [c["subseq_live"] for c in G["pairs"]]
    c = G["pairs"][0]
    for red, exp in zip((False, True), c["sim_seqs"]):
        assert mm2.sim_seqs(c["seq1"], mm2.b_read(Read("x", c["seq2"]), red, True, False)) == exp


def test_batches_against_reference_one_launch():
    from breakmer_b200 import sv_assembly_mm2 as mm2
    res = mm2.dedup_batches([as_batch(b["reads"]) for b in G["batches"]])
    for b, r in zip(G["batches"], res):
        assert unpack(r) == (b["checks"], b["kept"], b["redundant"], b["deleted"])


def test_incremental_read_batch_against_reference():
    from breakmer_b200 import sv_assembly_mm2 as mm2
    for b in G["batches"][:6]:
        reads = as_batch(b["reads"])
        rb = mm2.read_batch(reads[0][0], reads[0][1])
        checks = [True] + [rb.check_mer_read(p, r) for r, p in reads[1:]]
        assert checks == b["checks"]
        assert [x.read.id for x in rb.batch_reads] == b["kept"]
        assert [x.read.id for x in rb.batch_reads if x.redundant] == b["redundant"]
        assert sorted(rb.delete) == b["deleted"]


def test_random_batches_against_oracle():
    from breakmer_b200 import sv_assembly_mm2 as mm2
    rng = random.Random(4242)
    batches = [random_batch(rng, rng.randint(1, 60), "g%d" % t) for t in range(120)]
    for frac in (R.SUBSEQ_FRAC_MM2, R.SUBSEQ_FRAC_LIVE):
        res = mm2.dedup_batches([as_batch(b) for b in batches], frac=frac)
        for b, r in zip(batches, res):
            exp = R.dedup_batch(b, frac)
            assert unpack(r) == (exp[0], exp[1], exp[2], exp[3])


def test_errors():
    from breakmer_b200 import _lib, get_handle, sv_assembly_mm2 as mm2
    assert mm2.dedup_batches([]) == []
    one = mm2.dedup_batches([[(Read("a", "ACGT"), 0)]])[0]
    assert unpack(one) == ([True], ["a"], [], [])
    with pytest.raises(NameError):
        mm2.dedup_batches([[(Read("a", "ACGT"), 0), (Read("b", ""), 1)]])
    with pytest.raises(_lib.BreakmerError):
        get_handle(0).dedup_reads(["ACGT", "ACGT"], [0, 1], [0, 1], 0.9)      # batch_off does not cover the reads
    with pytest.raises(_lib.BreakmerError):
        get_handle(0).dedup_reads(["ACGT", "ACGT"], [0, 1], [0, 2], 1.5)      # threshold out of range


def test_batch_with_reads_above_4095_bases():
    """the alignments behind the decision chain go through bk_nw_batch, which has no length limit (olc.py:40-52): a batch
    mixing 4,200-4,400-base sequences (a duplicate, a contained one, a shifted one, an unrelated one) with short reads"""
    from breakmer_b200 import sv_assembly_mm2 as mm2
    rng = random.Random(77)
    g = "".join(rng.choice("ACGT") for _ in range(9000))
    batch = [("L0", g[:4300], 0), ("L1", g[:4300], 0), ("L2", g[40:4240], 0), ("s0", g[100:200], 0),
             ("L3", g[60:4400], 0), ("L4", g[4500:8800], 0), ("s1", g[100:200], 0), ("L5", g[4500:8800], 0)]
    for frac in (R.SUBSEQ_FRAC_MM2, R.SUBSEQ_FRAC_LIVE):
        got = mm2.dedup_batches([as_batch(batch)], frac=frac)[0]
        exp = R.dedup_batch(batch, frac)
        assert unpack(got) == (exp[0], exp[1], exp[2], exp[3])
    assert not all(R.dedup_batch(batch)[0])            # (some read was found redundant, i.e. the chain did something)
