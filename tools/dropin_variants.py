"""Drop-in throughput for chunking / pipelining variants of sv_processor.compare_kmers_batch (experiments)."""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.dropin_profile import Target
from breakmer_b200 import sv_processor, synth

regions = list(synth.config_regions("C2", 500))
d = tempfile.mkdtemp(prefix="bk_dropin_", dir="/dev/shm")
targets = [Target(r, d) for r in regions]
variants = [("160/chunk, 3 in flight (default)", dict()),
            ("160/chunk, 4 in flight", dict(inflight=4)),
            ("125/chunk, 4 in flight", dict(max_targets=125, inflight=4)),
            ("100/chunk, 5 in flight", dict(max_targets=100, inflight=5)),
            ("84/chunk, 6 in flight", dict(max_targets=84, inflight=6)),
            ("64/chunk, 8 in flight", dict(max_targets=64, inflight=8)),
            ("500 in one call", dict(max_targets=500)),
            ("apply thread, 160/chunk, 3 in flight", dict(apply_thread=True)),
            ("apply thread, 125/chunk, 4 in flight", dict(apply_thread=True, max_targets=125, inflight=4)),
            ("apply thread, 100/chunk, 5 in flight", dict(apply_thread=True, max_targets=100, inflight=5)),
            ("apply thread, 64/chunk, 8 in flight", dict(apply_thread=True, max_targets=64, inflight=8)),
            ("apply thread, 50/chunk, 6 in flight", dict(apply_thread=True, max_targets=50, inflight=6))]
only = sys.argv[1:]
for label, kw in variants:
    if only and not any(o in label for o in only):
        continue
    best = 1e9
    for rep in range(4):
        for t in targets:
            t.reset()
        t0 = time.time()
        sv_processor.compare_kmers_batch(targets, ingest="native", **kw)
        best = min(best, time.time() - t0)
    n_ctg = sum(len(t.kmers["clusters"]) for t in targets)
    print("%-42s %6.1f ms  %6.0f targets/s  (%d contigs)" % (label, 1e3 * best, 500 / best, n_ctg), flush=True)
