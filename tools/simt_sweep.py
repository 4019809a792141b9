"""Sweep of assemble_kernel on the host SIMT emulator (tests/sim) over regions of a BASELINE config, against the oracle.
    SIMT_ORDER=random:3 python tools/simt_sweep.py C2 0 80     # config, first region, one past the last"""
import sys, time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
import sim_util
from breakmer_b200 import synth
from oracle import assembler_py
from oracle.make_golden import oracle_sample_only
wl, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
bad = 0
t0 = time.time()
for i in range(a, b):
    region = synth.config_region(wl, i)
    _r, _c, _s, only = oracle_sample_only(region)
    exp = assembler_py.init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len)
    got, gst = sim_util.sim_init_assembly(only, region.reads, region.k, region.rc_thresh, region.read_len, simt=True)
    if got != exp:
        bad += 1
        print("MISMATCH", wl, i, flush=True)
print(wl, a, b, "bad", bad, "%.0fs" % (time.time() - t0))
