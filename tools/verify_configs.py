"""Full-size parity run: every region of a BASELINE.json configuration through the GPU
pipeline, compared with the oracle (C-accelerated nw) region by region.
Usage: python tools/verify_configs.py C1 C2 C3 C4 C5[:n]"""
import json
import multiprocessing as mp
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from breakmer_b200 import _lib, batch, synth          # noqa: E402
from oracle import assembler_py                      # noqa: E402
from oracle.make_golden import digest, oracle_sample_only   # noqa: E402

FULL = {"C1": 1, "C2": 500, "C3": 500, "C4": 100, "C5": 20000}


def oracle_one(args):
    cfg, i = args
    r = synth.config_region(cfg, i)
    _a, _b, _c, only = oracle_sample_only(r)
    ctg = assembler_py.init_assembly(only, r.reads, r.k, r.rc_thresh, r.read_len)
    return i, digest(sorted(only.items())), digest(ctg), len(only), len(ctg)


def main():
    h = _lib.Handle(0)
    report = {}
    for spec in sys.argv[1:]:
        cfg, _, n = spec.partition(":")
        n = int(n) if n else FULL[cfg]
        t0 = time.time()
        with mp.get_context("fork").Pool(os.cpu_count()) as pool:
            exp = dict((r[0], r[1:]) for r in pool.imap_unordered(oracle_one, [(cfg, i) for i in range(n)], chunksize=8))
        t_or = time.time() - t0
        bad = 0
        n_only = n_ctg = 0
        gpu_ms = 0.0
        chunk = 2500
        for a in range(0, n, chunk):
            regions = [synth.config_region(cfg, i) for i in range(a, min(n, a + chunk))]
            out = batch.run(h, batch.PackedBatch(regions))
            gpu_ms += out.gpu_ms
            for j, r in enumerate(regions):
                so = digest(sorted(out.sample_only(j).items()))
                ct = digest(out.contig_records(j))
                e = exp[a + j]
                if so != e[0] or ct != e[1] or out.region_status[j] != 0:
                    bad += 1
                n_only += e[2]
                n_ctg += e[3]
        report[cfg] = {"regions": n, "mismatching_regions": bad, "sample_only_kmers": n_only, "contigs": n_ctg,
                       "gpu_ms": round(gpu_ms, 2), "oracle_s_all_cores": round(t_or, 1)}
        print(cfg, report[cfg], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "verify_configs.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
