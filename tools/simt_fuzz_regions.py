"""Random-scenario fuzz of the whole device path against the oracle: regions with random k (11-31), read length (36-250),
coverage (0-300x), error / N / indel rates, event type and size, allele fraction and spurious-read fraction, one region per
C-ABI call (batch.run), sample-only k-mers and every contig record compared with the oracle.  TEST TOOL.

On a B200:          python tools/simt_fuzz_regions.py <seed> <n_regions> [max regions per call, default 1] [normal]
Without a GPU:      BK_LIB=tests/sim/libbreakmer_simt_TESTONLY.so python tools/simt_fuzz_regions.py <seed> <n_regions>
                    (the library on the host SIMT emulator of tests/sim; `python tests/sim_util.py` builds it)
"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from breakmer_b200 import _lib, batch, synth                      # noqa: E402


def scenarios(seed0, n, with_normal=False):
    """-> (rng, {k: [(kwargs, Region), ...]})"""
    rng = random.Random(seed0)
    by_k = {}
    for it in range(n):
        k = rng.choice([11, 15, 15, 21, 25, 31])
        rl = rng.choice([36, 50, 76, 100, 100, 150, 250])
        if rl <= k + 4:
            rl = k + 20
        ev = rng.choice([("del", rng.randint(20, 600), None), ("ins", rng.randint(5, 80)), ("tdup", rng.randint(30, 300)),
                         ("inv", rng.randint(100, 500)), ("none",), ("trl",)])
        kw = dict(seed=seed0 * 1000 + it, L=rng.randint(400, 2500), cov=rng.choice([0, 3, 10, 40, 120, 300]), k=k,
                  e=rng.choice([0.0, 0.002, 0.01, 0.03, 0.06]), event=ev, vaf=rng.choice([1.0, 0.5, 0.15]), rl=rl,
                  n_rate=rng.choice([0.0, 0.001, 0.02]), indel_p=rng.choice([0.0, 0.3, 0.8]),
                  spurious_frac=rng.choice([0.0, 0.0, 0.05, 0.3]), rl_jitter=rng.choice([0, 0, 10, 30]))
        if with_normal and rng.random() < 0.6:                        # tumour / normal pair: K4 normal subtraction
            kw.update(germline=True, normal_cov=rng.choice([5, 30, 100]))   # (drawn after the other knobs: seeds stay comparable)
        by_k.setdefault(k, []).append((kw, synth.make_region("f%d_%d" % (seed0, it), **kw)))
    return rng, by_k


def run_calls(h, rng, by_k, max_call, oracle_region):
    """every region through the device path in calls of 1..max_call regions of one k -> (calls, contigs, [mismatch descriptions])"""
    n_calls, nctg, bad = 0, 0, []
    for _k, items in sorted(by_k.items()):
        while items:
            take = rng.randint(1, max_call)
            call, items = items[:take], items[take:]
            n_calls += 1
            exp = [oracle_region(r) for _kw, r in call]
            out = batch.run(h, batch.PackedBatch([r for _kw, r in call]))
            for i, ((kw, _r), (only, ctg)) in enumerate(zip(call, exp)):
                nctg += len(ctg)
                if not (out.region_status[i] == 0 and out.sample_only(i) == only and out.contig_records(i) == ctg):
                    bad.append("region %d of a call of %d: %r" % (i, len(call), kw))
    return n_calls, nctg, bad


def main(argv):
    from test_gpu_pipeline import oracle_region
    seed0, n = int(argv[0]), int(argv[1])
    max_call = int(argv[2]) if len(argv) > 2 else 1
    rng, by_k = scenarios(seed0, n, len(argv) > 3 and argv[3] == "normal")
    h = _lib.Handle(0)
    t0 = time.time()
    n_calls, nctg, bad = run_calls(h, rng, by_k, max_call, oracle_region)
    for b in bad:
        print("MISMATCH", b)
    print("seed", seed0, "regions", n, "calls", n_calls, "contigs", nctg, "mismatches", len(bad), "%.0fs" % (time.time() - t0))
    h.close()
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
