"""The oracle's assembler restatement against `init_assembly` of the reference
itself (tests/golden/assembly_golden.json, see oracle/make_golden.py)."""
import hashlib
import json

import pytest

from conftest import golden
from breakmer_b200 import synth
from oracle import assembler_py, kmers_py, nw_py
from oracle.make_golden import digest, region_inputs_digest, oracle_sample_only

CASES = golden("assembly_golden.json")["cases"]


def _event(kw):
    kw = dict(kw)
    kw["event"] = tuple(kw["event"])
    return kw


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference(case):
    region = synth.make_region(case["name"], **_event(case["kwargs"]))
    assert region_inputs_digest(region) == case["inputs_sha256"], "generator drifted"
    _ref, _case, _sc, only = oracle_sample_only(region)
    assert len(only) == case["n_sample_only"]
    got = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
    assert len(got) == case["n_contigs"]
    assert digest(got) == case["contigs_sha256"]
    if "contigs" in case:
        assert got == case["contigs"]


def test_pure_python_nw_gives_same_contigs():
    case = CASES[2]
    region = synth.make_region(case["name"], **_event(case["kwargs"]))
    _ref, _case, _sc, only = oracle_sample_only(region)
    got = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, nw=nw_py.nw)
    assert digest(got) == case["contigs_sha256"]


def test_no_mers_no_contigs():
    assert assembler_py.init_assembly({}, [("@a:1:1:1:1/1_0", "ACGT" * 10, "I" * 40, False)], 15, 2, 40) == []


def test_group_reads_order_and_multiplicity():
    recs = [("@i:1:1:1:%d/1_0" % i, s, "I" * len(s), i == 1) for i, s in enumerate(["AAC", "GGT", "AAC", "TTT", "GGT", "AAC"])]
    g = assembler_py.group_reads(recs)
    assert [u.seq for u in g] == ["AAC", "GGT", "TTT"]
    assert [u.nreads for u in g] == [3, 2, 1]
    assert [u.rep_id for u in g] == [recs[0][0], recs[1][0], recs[3][0]]
    assert [u.indel_only for u in g] == [False, True, False]
