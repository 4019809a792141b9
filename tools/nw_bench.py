"""Throughput and latency of the two DP kernels of nw.cuh through bk_nw_batch: the score pass + traceback (default) and
the packed-cell kernel (BK_NW_PACKED=1).  Usage: python tools/nw_bench.py [read_len] [contig_len]"""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from breakmer_b200 import _lib      # noqa: E402


def main():
    rl = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    cl = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    rng = random.Random(5)
    n_seq = 4000
    seqs, pa, pb = [], [], []
    for i in range(n_seq):
        g = "".join(rng.choice("ACGT") for _ in range(cl + rl))
        contig = g[:cl]
        a = rng.randint(cl - rl + 5, cl - 20)          # read overlaps the contig's end and extends it
        read = "".join(c if rng.random() > 0.01 else rng.choice("ACGT") for c in g[a:a + rl])
        seqs.append(read); seqs.append(contig)
    h = _lib.Handle(0)
    for mode in ("trace", "packed"):
        if mode == "packed":
            os.environ["BK_NW_PACKED"] = "1"
        else:
            os.environ.pop("BK_NW_PACKED", None)
        for n_pairs in (148 * 4, 148 * 32, 400000):
            pa = [2 * (i % n_seq) for i in range(n_pairs)]        # seq1 = read (columns), seq2 = contig (rows)
            pb = [2 * (i % n_seq) + 1 for i in range(n_pairs)]
            h.nw_batch(seqs, pa, pb)
            h.kernel_times_reset(True)
            reps = 3
            for _ in range(reps):
                h.nw_batch(seqs, pa, pb)
            ms = h.kernel_times()["nw_batch"][0] / reps
            print("%-6s pairs %7d  kernel %9.3f ms  %8.1f G cells/s  %.2f us per pair per warp-slot" %
                  (mode, n_pairs, ms, n_pairs * rl * cl / ms / 1e6, ms * 1e3 / max(1, n_pairs / (148 * 8 * 4))))
    h.close()


if __name__ == "__main__":
    main()
