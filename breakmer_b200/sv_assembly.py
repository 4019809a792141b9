"""Drop-in for the reference's sv_assembly.init_assembly, running on the GPU.

    init_assembly(mers, fq_recs, kmer_len, rc_thresh, read_len) -> [contig]

keeps the reference signature (sv_assembly.py:30).  The returned objects carry
what the downstream code reads from a reference contig (SURVEY.md section 3.5,
sv_processor.py:731-746, sv_caller.py:159,242,271-277,634):

    .reads            set of the representative fq_read of each unique read
    .kmers            list of (mer, pos, less_than_half, dist_half, order) tuples
    .kmer_locs / get_kmer_locs()
    .aseq.seq / get_contig_seq()
    .aseq.counts / get_contig_counts()  with .indel_only, .others,
                      get_counts(p1, p2, sv_type), get_total_reads()
    get_total_read_support(), get_contig_len(), .kmer_len

The reference mutates its inputs while it runs (`fq_recs` loses keys, reads get
`.used = True`); the caller drops both right after the call
(sv_processor.py:643) and nothing downstream looks at them, so those mutations
are not replayed.
"""
from . import batch, get_handle


class assembly_counts:
    """Result view of the reference class of the same name (sv_assembly.py:160-221)."""

    def __init__(self, indel_only, others):
        self.indel_only = list(indel_only)
        self.others = list(others)

    def get_counts(self, p1, p2, sv_type):                       # sv_assembly.py:167-176
        if sv_type == 'indel' or sv_type == 'rearr':
            if p1 == p2:
                return self.indel_only[p1] + self.others[p1]
            return [x + y for x, y in zip(self.indel_only[p1:p2], self.others[p1:p2])]
        if p1 == p2:
            return self.others[p1]
        return self.others[p1:p2]

    def get_total_reads(self):                                   # sv_assembly.py:178-179
        return max(self.indel_only) + max(self.others)


class assembly_seq:
    def __init__(self, seq, counts):
        self.seq = seq
        self.counts = counts


class contig:
    def __init__(self, rec, reads, kmer_len):
        self.reads = set(reads)
        self.aseq = assembly_seq(rec["seq"], assembly_counts(rec["indel_only"], rec["others"]))
        self.kmer_locs = list(rec["kmer_locs"])
        self.kmers = [tuple(t) for t in rec["kmers"]]
        self.kmer_len = kmer_len
        self.setup = True

    def get_total_read_support(self):
        return self.aseq.counts.get_total_reads()

    def get_contig_len(self):
        return len(self.aseq.seq)

    def get_kmer_locs(self):
        return self.kmer_locs

    def get_contig_seq(self):
        return self.aseq.seq

    def get_contig_counts(self):
        return self.aseq.counts


class _AssemblyInput:
    """Region-like view of one init_assembly call for batch.PackedBatch."""

    def __init__(self, name, fq_recs, kmer_len, rc_thresh):
        self.name = name
        self.k = int(kmer_len)
        self.rc_thresh = int(rc_thresh)
        self.ref_fwd = ""
        self.sc_records = []
        self.normal_reads = []
        self.reads = []
        self.objs = []
        for seq, group in fq_recs.items():                       # insertion order == fq_recs order (Q9 policy)
            for fr in group:
                self.reads.append((fr.id, fr.seq, fr.qual, bool(fr.indel_only)))
                self.objs.append(fr)


def init_assembly_batch(calls, device=0):
    """calls: [(mers, fq_recs, kmer_len, rc_thresh, read_len), ...] with one common
    kmer_len and rc_thresh -> list (per call) of contig lists.  One GPU launch."""
    import numpy as np
    if not calls:
        return []
    inputs = [_AssemblyInput("r%d" % i, c[1], c[2], c[3]) for i, c in enumerate(calls)]
    pk = batch.PackedBatch(inputs, rc_thresh=int(calls[0][3]))
    pk.set_mers([c[0] for c in calls])
    pk.read_len = np.array([int(c[4]) for c in calls] + [0], dtype=np.int32)
    out = batch.run(get_handle(device), pk)
    bad = [i for i, s in enumerate(out.region_status) if s != 0]
    if bad:
        raise RuntimeError("init_assembly: device capacity exceeded in call(s) %s (contig longer than 4095 bases)" % bad)
    objs = [o for inp in inputs for o in inp.objs]
    result = []
    for i, c in enumerate(calls):
        recs = out.contig_records(i)
        ctgs = []
        for j, rec in enumerate(recs):
            cidx = int(out.ctg_reg_off[i]) + j
            ro, nr = out.reads_off[cidx]
            reads = [objs[int(r)] for r in out.reads[ro:ro + nr]]
            ctgs.append(contig(rec, reads, int(c[2])))
        result.append(ctgs)
    return result


def init_assembly(mers, fq_recs, kmer_len, rc_thresh, read_len):
    if len(mers) == 0:                                           # sv_assembly.py:33-34
        return []
    return init_assembly_batch([(mers, fq_recs, kmer_len, rc_thresh, read_len)])[0]
