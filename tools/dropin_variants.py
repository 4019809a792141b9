"""Drop-in throughput for a few chunking / threading variants (experiments)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.dropin_profile import Target
from breakmer_b200 import sv_processor, synth

regions = list(synth.config_regions("C2", 500))
d = tempfile.mkdtemp(prefix="bk_dropin_", dir="/dev/shm")
targets = [Target(r, d) for r in regions]
for label, kw in [("1 thread, 160/chunk, inflight 3", dict()), ("1 thread, 260/chunk", dict(max_targets=260)),
                  ("1 thread, 500 in one call", dict(max_targets=500)),
                  ("2 threads on device 0, 130/chunk", dict(devices=[0, 0], max_targets=130)),
                  ("2 threads, 260/chunk", dict(devices=[0, 0], max_targets=260)),
                  ("3 threads, 170/chunk", dict(devices=[0, 0, 0], max_targets=170)),
                  ("4 threads, 125/chunk", dict(devices=[0, 0, 0, 0], max_targets=125))]:
    best = 1e9
    for rep in range(4):
        for t in targets:
            t.reset()
        t0 = time.time()
        sv_processor.compare_kmers_batch(targets, ingest="native", **kw)
        best = min(best, time.time() - t0)
    print("%-40s %6.1f ms  %6.0f targets/s" % (label, 1e3 * best, 500 / best))
