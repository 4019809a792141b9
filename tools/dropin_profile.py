"""Time the reference-shaped Python drop-in (breakmer_b200.sv_processor.compare_kmers_batch) on the C2 panel:
python marshalling vs native ingest, with and without the native contig hand-off.
    python tools/dropin_profile.py [n_targets]"""
import cProfile
import logging
import os
import pstats
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from collections import OrderedDict                       # noqa: E402

from breakmer_b200 import sv_processor, synth, utils      # noqa: E402


class Params:
    def __init__(self, k):
        self.k = k
        self.opts = {"jellyfish": "jellyfish"}

    def get_kmer_size(self):
        return self.k

    def get_sr_thresh(self, kind):
        return 2


class Target:
    """The slice of sv_processor.target that compare_kmers touches (after extract_bam_reads / clean_reads)."""

    def __init__(self, region, d):
        base = os.path.join(d, region.name)
        with open(base + "_forward_refseq.fa", "w") as f:
            f.write(">%s\n%s\n" % (region.name, region.ref_fwd))
        with open(base + "_sv_reads_cleaned_filtered.fastq", "w") as f:
            for rid, seq, qual, _io in region.reads:
                f.write("%s\n%s\n+\n%s\n" % (rid, seq, qual))
        with open(base + "_sv_sc_seqs.fa", "w") as f:
            for name, seq in region.sc_records:
                f.write(">%s\n%s\n" % (name, seq))
        self.name = region.name
        self.params = Params(region.k)
        self.files = {"target_ref_fn": [base + "_forward_refseq.fa"], "cleaned_fq": base + "_sv_reads_cleaned_filtered.fastq",
                      "sv_sc_unmapped_fa": base + "_sv_sc_seqs.fa"}
        self.paths = {"kmers": d, "contigs": os.path.join(d, region.name + "_contigs")}
        self.kmers = {}
        self.read_len = region.read_len
        self.logger = logging.getLogger("root")
        self.region = region
        self.reset()

    def reset(self):
        recs = OrderedDict()
        for rid, seq, qual, io in self.region.reads:
            recs.setdefault(seq, []).append(utils.fq_read(rid, seq, qual, io))
        self.cleaned_read_recs = recs


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 500
    regions = list(synth.config_regions("C2", n))
    root = "/dev/shm" if os.path.isdir("/dev/shm") else None
    d = tempfile.mkdtemp(prefix="bk_dropin_", dir=root)
    targets = [Target(r, d) for r in regions]
    modes = [("python marshalling", dict(ingest="python")), ("native ingest", dict(ingest="native")),
             ("native ingest + native contig files", dict(ingest="native", write_contigs=True))]
    for label, kw in modes:
        best = None
        for rep in range(3):
            for t in targets:
                t.reset()
            t0 = time.time()
            sv_processor.compare_kmers_batch(targets, **kw)
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
        n_ctg = sum(len(t.kmers["clusters"]) for t in targets)
        print("%-40s %7.1f ms per %d targets (%d contigs) = %.0f targets/s" % (label, 1e3 * best, n, n_ctg, n / best))
    for t in targets:
        t.reset()
    pr = cProfile.Profile()
    pr.enable()
    sv_processor.compare_kmers_batch(targets, ingest="native", write_contigs=True)
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(14)


if __name__ == "__main__":
    main()
