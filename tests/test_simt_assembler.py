"""assemble_kernel itself -- breakmer_b200/csrc/assemble.cuh + nw.cuh, the product source -- on the host SIMT emulator of
tests/sim/simt_host.h: one CTA of W warps x 32 lanes as fibers, barriers at every warp collective and __syncthreads(),
lanes run one at a time between barriers (the most adversarial interleaving for "lane 0 writes, every lane reads" code).
The single-lane build (tests/test_sim_assembler.py) checks the control logic; this one also checks what only exists with
32 lanes and several warps: lane-strided loops, ballots and shuffles, the round protocol between the controller and the
aligner warps, the warp DP kernels, and warp-level races (a lane that reads a flag after lane 0 has already changed it
shows up as lanes diverging into different collectives -- three such races were found and fixed with this test).
Parity of the compiled sm_100a kernel is tests/test_gpu_pipeline.py."""
import pytest

import sim_util
from breakmer_b200 import synth
from oracle import assembler_py
from oracle.make_golden import oracle_sample_only, region_scenarios

SCEN = region_scenarios()
PICK = [SCEN[i] for i in (0, 4, 9, 10, 14, 20, 25, 33, 41, 47, 59)] + SCEN[-3:]


def _expected(name, kw):
    region = synth.make_region(name, **kw)
    _r, _c, _s, only = oracle_sample_only(region)
    stats = {}
    exp = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, stats=stats)
    return region, only, exp, stats


@pytest.mark.parametrize("name,kw", PICK, ids=[p[0] for p in PICK])
def test_kernel_on_the_emulator_matches_the_oracle(name, kw):
    region, only, exp, stats = _expected(name, kw)
    got, gst = sim_util.sim_init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, simt=True)
    assert got == exp
    assert gst["check_align"] == stats.get("check_align", 0)
    assert gst["cells"] * 2 == stats.get("cells", 0)      # one sweep serves both olc.nw calls


@pytest.mark.parametrize("spec_w,score_table", [(1, True), (2, True), (8, True), (4, False)])
def test_other_widths_and_the_packed_cell_kernel(spec_w, score_table):
    for name, kw in (SCEN[3], SCEN[14], SCEN[-2]):
        region, only, exp, _stats = _expected(name, kw)
        got, _gst = sim_util.sim_init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, simt=True,
                                               spec_w=spec_w, score_table=score_table)
        assert got == exp, (name, spec_w, score_table)


@pytest.mark.parametrize("order", ["reverse", "random:7"])
def test_other_interleavings_and_full_size_regions(order, monkeypatch):
    """regions of the BASELINE configs (the panel's longest region, a tumour/normal one, a deep amplicon, exome regions)
    with the threads taking their turns in descending / pseudo-random order instead of ascending"""
    monkeypatch.setenv("SIMT_ORDER", order)
    for wl, i in (("C2", 76), ("C3", 5), ("C5", 2)) + ((("C4", 3),) if order == "reverse" else ()):
        region = synth.config_region(wl, i)
        _r, _c, _s, only = oracle_sample_only(region)
        exp = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
        got, _gst = sim_util.sim_init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, simt=True)
        assert got == exp, (wl, i, order)
