"""Experiment helper: N handles in flight (threads), phase-profile build, dumps a region timeline."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("BK_PHASE_PRINT", "1")
from breakmer_b200 import _lib, batch, synth
n_fly = int(sys.argv[1]); steps = int(sys.argv[2])
regions = list(synth.config_regions("C2", n=500))
pk = batch.PackedBatch(regions)
hs = [_lib.Handle(0) for _ in range(n_fly)]
for h in hs:
    batch.upload(h, pk)
    for _ in range(3):
        batch.run(h, pk, resident=True, decode=False)
print("==== timed", file=sys.stderr, flush=True)
os.environ["BK_TIMELINE_DUMP"] = "gpurun_out/timeline.txt"
def w(j):
    for s in range(j, steps, n_fly):
        batch.run(hs[j], pk, resident=True, decode=False)
t0 = time.time()
ts = [threading.Thread(target=w, args=(j,)) for j in range(n_fly)]
[t.start() for t in ts]; [t.join() for t in ts]
print("wall per step ms", 1000 * (time.time() - t0) / steps, file=sys.stderr)
