"""GPU parity on random scenarios (tools/simt_fuzz_regions.py): regions with random k (11-31), read length, coverage, error
/ N / indel rates, event, allele fraction, spurious reads, tumour / normal pairs, run in calls of 1-8 regions of very
different sizes, against the oracle.  The same generator was run over 3,100 regions on the emulated library
(profiles/r2_emulator_runs.md); this is the slice the B200 sees in `pytest -m gpu` (named to run last)."""
import os
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_random_scenarios_in_multi_region_calls():
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    try:
        import simt_fuzz_regions as fz
    finally:
        sys.path.pop(0)
    from breakmer_b200 import _lib
    from test_gpu_pipeline import oracle_region
    rng, by_k = fz.scenarios(32, 48, with_normal=True)
    assert len(by_k) >= 4                                   # several k in one run
    h = _lib.Handle(0)
    try:
        n_calls, n_contigs, bad = fz.run_calls(h, rng, by_k, 8, oracle_region)
    finally:
        h.close()
    assert not bad, bad
    assert n_calls >= 8 and n_contigs > 50
