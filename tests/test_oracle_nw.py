"""The oracle's `nw` restatements against the reference's own outputs
(tests/golden/nw_golden.json was produced by /root/reference/olc.py itself,
see oracle/make_golden.py)."""
import random

import pytest

from conftest import golden
from oracle import nw_py


@pytest.fixture(scope="module")
def cases():
    return golden("nw_golden.json")["cases"]


def test_reference_examples_known_answers(cases):
    # the three example pairs at olc.py:11-16, both ways
    got = [tuple(c["out"][2:]) for c in cases[:6]]
    assert got == [(101, 39, 62, 0, 56), (87, 0, 101, 39, 6),
                   (101, 33, 68, 0, 65), (96, 0, 101, 33, 9),
                   (95, 94, 0, 0, 0), (39, 33, 6, 0, 0)]


def test_python_restatement_matches_reference(cases):
    for c in cases:
        assert list(nw_py.nw(c["seq1"], c["seq2"])) == c["out"]


def test_c_restatement_matches_reference(cases):
    assert nw_py.c_lib() is not None
    for c in cases:
        assert list(nw_py.nw_fast(c["seq1"], c["seq2"])) == c["out"]


def test_gapless_alignment_is_a_substring(cases):
    # property the device path relies on: alignment strings minus '-' are the
    # aligned spans seq1[j0:m] / seq2[i0:prei]  (kernel A8)
    for c in cases:
        a1, a2, prej, j0, prei, i0, _ = c["out"]
        assert a1.replace("-", "") == c["seq1"][j0:prej]
        assert a2.replace("-", "") == c["seq2"][i0:prei]


def test_empty_sequence_raises_nameerror():
    with pytest.raises(NameError):
        nw_py.nw("", "ACGT")
    with pytest.raises(NameError):
        nw_py.nw_fast("ACGT", "")


def test_c_equals_python_on_fresh_random_pairs():
    rng = random.Random(5)
    for _ in range(150):
        a = "".join(rng.choice("ACGTN") for _ in range(rng.randint(1, 90)))
        b = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 90)))
        assert nw_py.nw(a, b) == nw_py.nw_fast(a, b)


def test_identity_threshold_is_an_integer_test():
    # Q27: round(s/span, 2) < 0.90  <=>  200*s < 179*span ; s < min_len/4.0 <=> 4*s < min_len
    for span in range(1, 1200):
        for s in range(0, span + 1):
            assert (round(float(s) / float(span), 2) < 0.90) == (200 * s < 179 * span), (s, span)
