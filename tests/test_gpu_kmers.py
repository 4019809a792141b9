"""GPU parity of the k-mer stage entry points (rows K1-K4) against the oracle."""
import random

import numpy as np
import pytest

from conftest import golden
from breakmer_b200 import synth
from oracle import kmers_py
from oracle.make_golden import oracle_sample_only

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from breakmer_b200 import _lib
    h = _lib.Handle(0)
    yield h
    h.close()


def as_dict(mers, counts, k):
    from breakmer_b200._lib import code_to_mer
    assert np.all(mers[1:] > mers[:-1]), "mers must be strictly ascending"
    return {code_to_mer(m, k): int(c) for m, c in zip(mers, counts)}


@pytest.mark.parametrize("k", [1, 4, 15, 21, 31])
def test_count_kmers_matches_oracle(handle, k):
    rng = random.Random(k)
    seqs = ["".join(rng.choice("ACGTNacgtn"[: 5 if i % 7 else 10]) for _ in range(rng.choice([0, 1, k - 1, k, k + 1, 60, 100, 1500])))
            for i in range(300)]
    mult = [rng.randint(1, 5) for _ in seqs]
    got = as_dict(*handle.count_kmers(seqs, k), k)
    assert got == kmers_py.count_kmers(seqs, k)
    got_m = as_dict(*handle.count_kmers(seqs, k, mult=mult), k)
    exp = {}
    for s, m in zip(seqs, mult):
        for mer, c in kmers_py.count_kmers([s], k).items():
            exp[mer] = exp.get(mer, 0) + c * m
    assert got_m == exp


def test_count_kmers_edge_cases(handle):
    assert len(handle.count_kmers([], 15)[0]) == 0
    assert len(handle.count_kmers(["", "ACG"], 15)[0]) == 0
    assert as_dict(*handle.count_kmers(["A" * 40], 15), 15) == {"A" * 15: 26}
    # windows never span records
    assert as_dict(*handle.count_kmers(["ACGT", "ACGT"], 4), 4) == {"ACGT": 2}


def test_large_single_record(handle):
    rng = random.Random(9)
    s = "".join(rng.choice("ACGT") for _ in range(300000))
    got = as_dict(*handle.count_kmers([s], 15), 15)
    assert got == kmers_py.count_kmers([s], 15)


def test_sample_only_matches_oracle_on_golden_regions(handle):
    from breakmer_b200._lib import mer_to_code
    for case in golden("kmers_golden.json")["cases"]:
        kw = dict(case["kwargs"]); kw["event"] = tuple(kw["event"])
        region = synth.make_region(case["name"], **kw)
        k = region.k
        ref = handle.count_kmers([region.ref_fwd, kmers_py.revcomp(region.ref_fwd)], k)
        cs = handle.count_kmers([r[1] for r in region.reads], k)
        sc = handle.count_kmers([r[1] for r in region.sc_records], k)
        oref, ocase, osc, only = oracle_sample_only(region)
        assert as_dict(*ref, k) == oref
        assert as_dict(*cs, k) == ocase
        assert as_dict(*sc, k) == osc
        got = as_dict(*handle.sample_only(k, cs, sc[0], ref[0]), k)
        assert got == only
        assert [[m, c] for m, c in sorted(got.items())] == case["sample_only"]


def test_normal_subtraction(handle):
    region = synth.config_region("C3", 3)
    k = region.k
    ref = handle.count_kmers([region.ref_fwd, kmers_py.revcomp(region.ref_fwd)], k)
    cs = handle.count_kmers([r[1] for r in region.reads], k)
    sc = handle.count_kmers([r[1] for r in region.sc_records], k)
    nm = handle.count_kmers([r[1] for r in region.normal_reads], k)
    _r, _c, _s, only = oracle_sample_only(region)
    _r, _c, _s, only_no_normal = kmers_py.sample_only(region.ref_fwd, [x[1] for x in region.reads],
                                                      [x[1] for x in region.sc_records], k)
    assert len(only) < len(only_no_normal)          # the germline indel is really subtracted
    assert as_dict(*handle.sample_only(k, cs, sc[0], ref[0], nm[0]), k) == only
