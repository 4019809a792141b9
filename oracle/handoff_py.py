"""CPU restatement of the contig hand-off writers (sv_processor.contig.setup,
sv_processor.py:749-782).  TEST INFRASTRUCTURE ONLY: the product path
is bk_write_contigs (breakmer_b200/csrc/ingest.cuh).  Pinned against the reference's own methods
by oracle/make_golden_ingest.py -> tests/golden/handoff_cases.json.

The reference iterates `self.reads`, a Python set, whose order is not defined; here (and in the
golden generator, which hands the reference an ordered container) reads keep the given order.
"""


def contig_files(contig_id, seq, kmers, reads):
    """{relative path: text} of one contig; reads = [(id, seq, qual)], kmers = [mer string, ...]."""
    return {
        "%s/%s.fq" % (contig_id, contig_id): "".join("%s\n%s\n+\n%s\n" % r for r in reads),      # :767-772
        "%s/%s.fa" % (contig_id, contig_id): ">contig1\n" + seq,                                   # :776-781
    }


def cluster_text(contig_id, kmers, reads):
    """write_cluster_file (:758-763)."""
    return "%s %d\n%s\n%s\n\n" % (contig_id, len(kmers), ",".join(kmers), ",".join(r[0] for r in reads))


def target_files(contigs):
    """All files of one target: contigs = [(seq, [mer, ...], [(id, seq, qual), ...]), ...] in acceptance order.
    Returns ({relative path: text}, cluster file text or None).  The cluster file is rewritten for every contig, so
    the last one survives (sv_processor.py:759 opens it with 'w')."""
    files = {}
    cluster = None
    for n, (seq, kmers, reads) in enumerate(contigs, 1):
        cid = "contig%d" % n                                                                      # :653
        files.update(contig_files(cid, seq, kmers, reads))
        cluster = cluster_text(cid, kmers, reads)
    return files, cluster
