"""Experiment helper: run one resident pass of a workload and let the library print
its per-phase cycle counters (needs the -DBK_PHASE_PROF build, BK_LIB=...)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BK_PHASE_PRINT", "1")
from breakmer_b200 import _lib, batch, synth

wl = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 500
regions = list(synth.config_regions(wl, n=n))
pk = batch.PackedBatch(regions)
h = _lib.Handle(0)
batch.upload(h, pk)
for i in range(2):
    res = batch.run(h, pk, resident=True, decode=False)
    print("gpu_ms", res.gpu_ms, "check_align", res.n_check_align, "cells", res.n_dp_cells, file=sys.stderr)
