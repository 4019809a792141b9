"""K-mer stage oracle (rows K1-K4).  Jellyfish is absent, so these are
oracle-derived known answers ("parity unpinned", oracle/kmers_py.py) plus the
semantics the restatement claims."""
from conftest import golden
from breakmer_b200 import synth
from oracle import kmers_py
from oracle.make_golden import digest, region_inputs_digest, oracle_sample_only


def test_counts_are_strand_specific_occurrences():
    c = kmers_py.count_kmers(["ACGTACGT", "ACGTA"], 4)
    assert c == {"ACGT": 3, "CGTA": 2, "GTAC": 1, "TACG": 1}


def test_windows_with_non_acgt_are_skipped_and_case_folded():
    c = kmers_py.count_kmers(["ACNGTAC", "acgt"], 3)
    assert c == {"GTA": 1, "TAC": 1, "ACG": 1, "CGT": 1}


def test_record_shorter_than_k_contributes_nothing():
    assert kmers_py.count_kmers(["ACG", ""], 4) == {}


def test_sample_only_algebra():
    ref = "AAAACCCCGGGGTTTT"
    reads = ["CCCCGGGGAT", "GGGGATCA", "GGGGATCA"]
    sc = ["GGATCA"]
    r, case, csc, only = kmers_py.sample_only(ref, reads, sc, 4)
    assert set(only) == (set(case) & set(csc)) - set(r)
    assert only["GATC"] == 2 and only["GGAT"] == 3
    _r, _c, _s, only_n = kmers_py.sample_only(ref, reads, sc, 4, normal_seqs=["TGGATT"])
    assert "GGAT" not in only_n and "GATC" in only_n


def test_known_answers():
    for case in golden("kmers_golden.json")["cases"]:
        kw = dict(case["kwargs"])
        kw["event"] = tuple(kw["event"])
        region = synth.make_region(case["name"], **kw)
        assert region_inputs_digest(region) == case["inputs_sha256"]
        ref, cs, sc, only = oracle_sample_only(region)
        assert (len(ref), len(cs), len(sc)) == (case["n_ref"], case["n_case"], case["n_sc"])
        assert digest(sorted(ref.items())) == case["ref_sha256"]
        assert digest(sorted(cs.items())) == case["case_sha256"]
        assert [list(x) for x in sorted(only.items())] == case["sample_only"]


def test_vectorised_mer_decoding_matches_scalar():
    import random
    import numpy as np
    from breakmer_b200 import _lib
    rng = random.Random(5)
    for k in (1, 2, 11, 15, 21, 25, 31):
        codes = np.array([rng.getrandbits(2 * k) for _ in range(300)] + [0, (1 << (2 * k)) - 1], dtype=np.uint64)
        assert _lib.codes_to_mers(codes, k) == [_lib.code_to_mer(c, k) for c in codes]
        assert [_lib.mer_to_code(m) for m in _lib.codes_to_mers(codes, k)] == codes.tolist()
    assert _lib.codes_to_mers(np.zeros(0, np.uint64), 15) == []
