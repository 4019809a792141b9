"""Where a drop-in call spends its time (experiments): wraps the steps of sv_processor.compare_kmers_batch with timers."""
import os, sys, tempfile, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.dropin_profile import Target
from breakmer_b200 import sv_processor, synth, batch, ingest, shard

acc = collections.defaultdict(float)
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **kw):
        t0 = time.perf_counter()
        try:
            return f(*a, **kw)
        finally:
            acc[label] += time.perf_counter() - t0
    setattr(obj, name, g)

wrap(ingest.Ingest, "files", "ingest.files")
wrap(ingest.Ingest, "write_sample_kmers", "write_sample_kmers")
wrap(batch, "submit", "submit")
wrap(batch, "wait", "wait")
wrap(batch, "run", "run")
wrap(batch, "BatchOutput", "BatchOutput")
wrap(sv_processor, "_LazyReads", "_LazyReads")
wrap(sv_processor, "_apply_chunk", "_apply_chunk(total)")
wrap(sv_processor, "_pack_targets", "_pack_targets(total)")

regions = list(synth.config_regions("C2", 500))
d = tempfile.mkdtemp(prefix="bk_dropin_", dir="/dev/shm")
targets = [Target(r, d) for r in regions]
for kw in (dict(), dict(max_targets=500)):
    for rep in range(4):
        for t in targets:
            t.reset()
        acc.clear()
        t0 = time.perf_counter()
        sv_processor.compare_kmers_batch(targets, ingest="native", **kw)
        tot = time.perf_counter() - t0
    print(kw, "total %.1f ms" % (1e3 * tot), {k: round(1e3 * v, 1) for k, v in sorted(acc.items())})
