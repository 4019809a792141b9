"""Drop-in for `target.compare_kmers()` (sv_processor.py:609-645) and a batched
form for the region loop (sv_processor.py:185-201).

`compare_kmers(target)` takes the reference's own `target` object (or anything
shaped like it) after `set_ref_data`, `extract_bam_reads` and `clean_reads` have
run, and leaves it in the state the reference method leaves it in:

    target.kmers['clusters']   list of contigs (breakmer_b200.sv_assembly.contig)
    target.kmers['ref'|'case'|'case_sc'|'case_only'] = {}       (:634-636, :644)
    target.files['sample_kmers']  "<kmers path>/<name>_sample_kmers.out" written
                                  as "<mer>\\t<case count>" lines (:625-632)
    target.files['kmer_clusters'] set (:639)
    target.cleaned_read_recs = None                              (:643)

What it reads from the target: files['target_ref_fn'][0] (forward reference
FASTA; the reverse-complement file the reference also counts is by construction
its reverse complement, utils.py:367-371, and is derived on the device),
files['cleaned_fq'], files['sv_sc_unmapped_fa'], cleaned_read_recs, read_len,
paths['kmers'], name, params.get_kmer_size(), params.get_sr_thresh('min').
An optional files['normal_fq'] enables normal-sample subtraction (K4).

`compare_kmers_batch(targets)` does the same for many targets in ONE device
pass; the reference loop becomes: extract+clean all targets, one batched call,
then resolve_sv per target.

With `ingest="native"` the inputs are not marshalled from the Python objects at all:
the four files of every target (target_ref_fn[0], cleaned_fq -- the filtered FASTQ
get_fastq_reads wrote next to cleaned_read_recs, utils.py:206,237 -- sv_sc_unmapped_fa,
normal_fq) are parsed by the library's host threads straight into page-locked memory
(bk_ingest_files).  The reads of the returned contigs are then
fresh fq_read objects built from the parsed records (same .id/.seq/.qual/.indel_only as the
caller's; of these the rest of the reference only reads .id/.seq/.qual).
`write_contigs=True` additionally writes, for every contig of every target, the files
contig.setup writes before blat (sv_processor.py:749-782) under target.paths['contigs'], on the
library's host threads (bk_write_contigs).
"""
import os

from . import batch, get_handle, utils
from .sv_assembly import contig


class _TargetInput:
    def __init__(self, trgt):
        self.name = trgt.name
        self.k = int(trgt.params.get_kmer_size())
        self.rc_thresh = int(trgt.params.get_sr_thresh('min'))
        refs = utils.read_sequences(trgt.files['target_ref_fn'][0])
        self.ref_fwd = refs[0] if refs else ""
        self.reads = []
        self.objs = []
        for seq, group in trgt.cleaned_read_recs.items():
            for fr in group:
                self.reads.append((fr.id, fr.seq, fr.qual, bool(fr.indel_only)))
                self.objs.append(fr)
        self.sc_records = [("sc", s) for s in utils.read_sequences(trgt.files['sv_sc_unmapped_fa'])]
        nfq = trgt.files.get('normal_fq') if hasattr(trgt.files, "get") else None
        self.normal_reads = [("n", s) for s in utils.read_sequences(nfq)] if nfq else []
        self.read_len = int(trgt.read_len)


class _LazyReads:
    """objs[i] for the native ingest: fq_read of record i, built on first use."""

    def __init__(self, pk):
        self._pk = pk
        self._cache = {}
        self._cols = None

    def __getitem__(self, i):
        fr = self._cache.get(i)
        if fr is None:
            if self._cols is None:
                self._cols = (self._pk.read_ids, self._pk.read_seqs(), self._pk.read_quals())
            ids, seqs, quals = self._cols
            fr = self._cache[i] = utils.fq_read(ids[i], seqs[i], quals[i], bool(self._pk.read_flags[i]))
        return fr


_ingests = {}


def _get_ingest():
    import threading
    from . import ingest as _ingest
    key = threading.get_ident()                    # an ingest object's buffer is reused call to call: one per thread
    if key not in _ingests:
        _ingests[key] = _ingest.Ingest()
    return _ingests[key]


class _K:
    def __init__(self, k):
        self.k = k


def compare_kmers_batch(targets, device=0, ingest="python", write_contigs=False):
    import numpy as np
    if not targets:
        return
    if ingest == "native":
        get = lambda t, key: (t.files.get(key) if hasattr(t.files, "get") else None)   # noqa: E731
        k = int(targets[0].params.get_kmer_size())
        if any(int(t.params.get_kmer_size()) != k for t in targets):
            raise ValueError("one k per batch")
        nfq = [get(t, 'normal_fq') for t in targets]
        pk = _get_ingest().files([t.files['target_ref_fn'][0] for t in targets], [t.files['cleaned_fq'] for t in targets],
                                 [t.files['sv_sc_unmapped_fa'] for t in targets], normal=nfq if any(nfq) else None,
                                 k=k, rc_thresh=int(targets[0].params.get_sr_thresh('min')),
                                 names=[t.name for t in targets])
        for i, t in enumerate(targets):             # target.read_len is what get_fastq_reads returned (utils.py:236,246)
            pk.read_len[i] = int(t.read_len)
        inputs = [_K(k)] * len(targets)
        objs = _LazyReads(pk)
    elif ingest == "python":
        inputs = [_TargetInput(t) for t in targets]
        pk = batch.PackedBatch(inputs, rc_thresh=inputs[0].rc_thresh)
        pk.read_len = np.array([inp.read_len for inp in inputs] + [0], dtype=np.int32)
        objs = [o for inp in inputs for o in inp.objs]
    else:
        raise ValueError("ingest must be 'python' or 'native'")
    if write_contigs and ingest != "native":
        raise ValueError("write_contigs needs ingest='native' (the writer reads the parsed record text)")
    res = batch.run(get_handle(device), pk, decode=False)
    if write_contigs:
        # contig.setup's files for every contig of every target in one pass (sv_processor.py:749-782, bk_write_contigs);
        # the reference-side contig.__init__ then skips its own setup() call
        _get_ingest().write_contigs(res, pk, [t.paths['contigs'] for t in targets],
                                    [os.path.join(t.paths['kmers'], t.name + "_sample_kmers_merged.out") for t in targets])
    out = batch.BatchOutput(res, pk)
    for trgt in targets:
        trgt.files['sample_kmers'] = os.path.join(trgt.paths['kmers'], trgt.name + "_sample_kmers.out")
    if ingest == "native":
        # the "<mer>\t<count>" files of all targets in one multi-threaded sweep (bk_write_sample_kmers)
        _get_ingest().write_sample_kmers(res, pk.k, [t.files['sample_kmers'] for t in targets])
    for i, trgt in enumerate(targets):
        if out.region_status[i] != 0:
            raise RuntimeError("compare_kmers: device capacity exceeded for target %s" % trgt.name)
        n_only = int(out.so_off[i + 1] - out.so_off[i])
        if ingest != "native":
            with open(trgt.files['sample_kmers'], 'w') as f:
                for mer, cnt in out.sample_only(i).items():
                    f.write("\t".join([mer, str(cnt)]) + "\n")
        for key in ('ref', 'case', 'case_sc'):
            trgt.kmers[key] = {}
        logger = getattr(trgt, "logger", None)
        if logger is not None:
            logger.info('Writing %d sample-only kmers to file %s' % (n_only, trgt.files['sample_kmers']))
        trgt.files['kmer_clusters'] = os.path.join(trgt.paths['kmers'], trgt.name + "_sample_kmers_merged.out")
        ctgs = []
        for j, rec in enumerate(out.contig_records(i, with_reads=False)):
            cidx = int(out.ctg_reg_off[i]) + j
            ro, nr = out.reads_off[cidx]
            ctgs.append(contig(rec, [objs[int(r)] for r in out.reads[ro:ro + nr]], inputs[i].k))
        trgt.kmers['clusters'] = ctgs
        trgt.cleaned_read_recs = None
        trgt.kmers['case_only'] = {}


def compare_kmers(target, device=0, ingest="python"):
    compare_kmers_batch([target], device=device, ingest=ingest)
