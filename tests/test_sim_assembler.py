"""Control logic of the device assembler (breakmer_b200/csrc/assemble.cuh) checked
on the CPU through tests/sim: the same source compiled by g++ with a one-lane
"warp" and a scalar DP.  This is a debugging aid for the GPU-less build container;
the GPU parity tests (tests/test_gpu_pipeline.py) are the ones that exercise the
product kernels."""
import pytest

import sim_util
from breakmer_b200 import synth
from oracle import assembler_py
from oracle.make_golden import oracle_sample_only, region_scenarios

SCEN = region_scenarios()
PICK = [SCEN[i] for i in (0, 3, 4, 9, 14, 20, 23, 25, 29, 33, 41, 47, 53, 59)] + SCEN[-6:]


@pytest.mark.parametrize("name,kw", PICK, ids=[p[0] for p in PICK])
def test_sim_matches_oracle(name, kw):
    region = synth.make_region(name, **kw)
    _r, _c, _s, only = oracle_sample_only(region)
    stats = {}
    exp = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, stats=stats)
    got, gst = sim_util.sim_init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
    assert got == exp
    assert gst["check_align"] == stats.get("check_align", 0)
    assert gst["cells"] * 2 == stats.get("cells", 0)      # one sweep serves both olc.nw calls
