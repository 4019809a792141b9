"""Randomised GPU-vs-oracle parity run: N random regions (event type, coverage, error rate, k, read length, jitter,
spurious reads, indel_only mix), batched by k, compared region by region with the oracle.
Usage: python tools/fuzz_gpu.py N [seed]"""
import multiprocessing as mp
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from breakmer_b200 import _lib, batch, synth          # noqa: E402
from oracle import assembler_py                      # noqa: E402
from oracle.make_golden import digest, oracle_sample_only   # noqa: E402


def params(t, seed):
    rng = random.Random(seed * 1000003 + t)
    ev = rng.choice([("del", rng.randint(20, 800), None), ("ins", rng.randint(10, 60)), ("inv", rng.randint(80, 600)),
                     ("tdup", rng.randint(60, 400)), ("trl",), ("none",)])
    return dict(seed=300000 + seed * 100000 + t, L=rng.randint(400, 6000), cov=rng.choice([40, 100, 200, 500, 1000]),
                k=rng.choice([11, 15, 15, 21, 25, 31]), e=rng.choice([0, 0.002, 0.01, 0.03]), event=ev,
                vaf=rng.choice([1.0, 0.5, 0.25]), indel_p=rng.choice([0, 0.3, 1.0]),
                rl=rng.choice([60, 75, 100, 100, 150, 250, 300]), rl_jitter=rng.choice([0, 0, 10, 30]),
                spurious_frac=rng.choice([0, 0, 0.01, 0.03]))


def oracle_one(args):
    t, seed = args
    r = synth.make_region("z%d" % t, **params(t, seed))
    _a, _b, _c, only = oracle_sample_only(r)
    ctg = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
    return t, digest(sorted(only.items())), digest(ctg)


def main():
    n = int(sys.argv[1])
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        exp = dict((r[0], r[1:]) for r in pool.imap_unordered(oracle_one, [(t, seed) for t in range(n)], chunksize=4))
    by_k = {}
    for t in range(n):
        by_k.setdefault(params(t, seed)["k"], []).append(t)
    h = _lib.Handle(0)
    bad = []
    for k, ts in sorted(by_k.items()):
        for w in (4, 1):
            h.set_option("spec_width", w)
            regions = [synth.make_region("z%d" % t, **params(t, seed)) for t in ts]
            out = batch.run(h, batch.PackedBatch(regions))
            for j, t in enumerate(ts):
                so = digest(sorted(out.sample_only(j).items()))
                ct = digest(out.contig_records(j))
                if (so, ct) != exp[t] or out.region_status[j] != 0:
                    bad.append((t, k, w, params(t, seed)))
        print("k=%d: %d regions checked at widths 4 and 1" % (k, len(ts)), flush=True)
    print("fuzz: %d regions, %d mismatches" % (n, len(bad)))
    for b in bad[:10]:
        print("MISMATCH", b)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
