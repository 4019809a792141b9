"""Oracle restatement of the reference greedy k-mer assembler.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates the *live* path of /root/reference/sv_assembly.py (SURVEY.md section
3.3 and rows A1-A13, T1), quirks included (Q5-Q30).  Dead code
(sv_assembly.py:67-98, `check_mer_read`'s disabled branches, `redundant`,
`mer_pos_d`) is not restated.  Line citations below are into
/root/reference/sv_assembly.py unless a file is named.

The restatement is index based (reads are numbered in `fq_recs` order) rather
than object based, so the same description doubles as the specification of the
device state machine; it is pinned against the reference itself by
tests/golden/assembly_golden.json (oracle/make_golden.py via oracle/ref_shim.py).

Order policy (the two places where the reference's
result depends on CPython-2 hash order, which cannot be observed):
  * Q9  `fq_recs.items()` order == insertion order == first occurrence in the
        cleaned FASTQ; deleting a key keeps the order of the others.
  * Q13 `check_alt_reads` walks its candidate set in ascending mer order.
"""
from . import nw_py


class UniqueRead:
    """One key of `fq_recs` (utils.py:239-244, Q28): all records with the same
    sequence; `rep` (= reads[0]) represents the group."""
    __slots__ = ("idx", "seq", "ids", "nreads", "indel_only", "used", "deleted")

    def __init__(self, idx, seq, rid, indel_only):
        self.idx = idx
        self.seq = seq
        self.ids = [rid]
        self.nreads = 1
        self.indel_only = bool(indel_only)   # of reads[0]
        self.used = False
        self.deleted = False

    @property
    def rep_id(self):
        return self.ids[0]


def group_reads(records):
    """records: iterable of (id, seq, qual, indel_only) in cleaned-FASTQ order.
    Returns the list of UniqueRead in fq_recs insertion order."""
    by_seq = {}
    out = []
    for rec in records:
        rid, seq, io = rec[0], rec[1], rec[3]
        u = by_seq.get(seq)
        if u is None:
            u = UniqueRead(len(out), seq, rid, io)
            by_seq[seq] = u
            out.append(u)
        else:
            u.ids.append(rid)
            u.nreads += 1
    return out


FOR, REV, MID = "for", "rev", "mid"


def read_kmers_ordered(seq, k, live, order):
    """get_read_kmers_ordered (:126-143).  Skips the last window (Q8); m uses
    floor division (Q26)."""
    m = len(seq) // 2
    out = [(seq[x:x + k], x, int(x < m), abs(x - m), order)
           for x in range(0, len(seq) - k) if seq[x:x + k] in live]
    if order == REV:
        out.reverse()
    elif order == MID:
        out.sort(key=lambda t: (t[2], t[3]))
    return out


def read_kmer_set(seq, k, live):
    """get_read_kmers (:147-155), also without the last window (Q8)."""
    return {seq[x:x + k] for x in range(0, len(seq) - k)} & live


class Contig:
    def __init__(self, seed_mer, read, k):
        # contig.__init__ (:417-426), assembly_seq/assembly_counts (:161-165,227-230)
        self.k = k
        self.seq = read.seq
        n = len(read.seq)
        self.indel_only = [0] * n
        self.others = [0] * n
        self._set_counts(0, n, read.nreads, read.indel_only)
        self.kmers = []
        self.kmer_locs = []
        self.checked = [seed_mer]            # Q24
        self.buffer = {read.idx}
        self.reads = set()
        self.setup = False
        self.batch = [(read, True)]          # read_batch.batch_reads as (read, aligned)
        self.alt = []
        self.delete = []

    # ---- assembly_counts (:160-221) -------------------------------------
    def _set_counts(self, start, end, nreads, indel_only):            # :195-199
        vec = self.indel_only if indel_only else self.others
        vec[start:end] = [x + nreads for x in vec[start:end]]

    def _set_superseq(self, read, start, end):                         # :181-193 (Q17)
        n = len(read.seq)
        t_io = [read.nreads if read.indel_only else 0] * n
        t_ot = [0 if read.indel_only else read.nreads] * n
        t_io[start:end] = [x + y for x, y in zip(t_io[start:end], self.indel_only)]
        t_ot[start:end] = [x + y for x, y in zip(t_ot[start:end], self.others)]
        self.indel_only, self.others = t_io, t_ot
        self.seq = read.seq                                            # :234

    def _extend_counts(self, l, nreads, indel_only, post):            # :201-221
        ext, fill = [nreads] * l, [0] * l
        a, b = (ext, fill) if indel_only else (fill, ext)
        if post:
            self.indel_only = self.indel_only + a
            self.others = self.others + b
        else:
            self.indel_only = a + self.indel_only
            self.others = b + self.others

    def total_reads(self):                                             # :178-179
        return max(self.indel_only) + max(self.others)

    # ---- k-mer bookkeeping ------------------------------------------------
    def set_kmers(self, live):                                         # :548-550
        self.setup = True
        self.kmers = read_kmers_ordered(self.seq, self.k, live, MID)

    def set_kmer_locs(self):                                           # :434-438 (Q25)
        locs = [0] * len(self.seq)
        for t in self.kmers:
            p = self.seq.find(t[0])
            locs[p:p + self.k] = [x + 1 for x in locs[p:p + self.k]]
        self.kmer_locs = locs

    # ---- alignment decision tree (:449-546) ----------------------------------
    def check_align(self, read, mer, live, grow, nw):
        C, R = self.seq, read.seq
        v1 = nw(C, R)
        v2 = nw(R, C)
        lc, lr = len(C), len(R)
        # :459-464 in integers (Q27): round(s/span,2) < 0.90  <=>  200*s < 179*span
        s1, s2 = v1[6], v2[6]
        bad1 = 4 * s1 < min(lc, lr) or 200 * s1 < 179 * (v1[2] - v1[3])
        bad2 = 4 * s2 < min(lc, lr) or 200 * s2 < 179 * (v2[2] - v2[3])
        if bad1 and bad2:
            return False
        if s1 == s2 and v1[3] == 0 and v1[5] == 0 and lc == lr:        # :466 (Q16)
            return True
        if s1 == s2:
            if lc < lr or (v1[2] == lc and v1[3] == 0):                # :471
                self._set_superseq(read, v1[5], v1[4])
                if grow:
                    self.set_kmers(live)
                return True
            if lr < lc or (v2[2] == lr and v2[3] == 0):                # :480
                self._set_counts(v2[5], v2[4], read.nreads, read.indel_only)
                return True
            i11 = v1[0].replace('-', '').find(mer)                     # :485-496
            i12 = v1[1].replace('-', '').find(mer)
            i21 = v2[0].replace('-', '').find(mer)
            i22 = v2[1].replace('-', '').find(mer)
            if i11 > -1 and i12 > -1:
                if (i21 == -1 and i22 == -1) or abs(i21 - i22) > abs(i11 - i12):
                    self._contig_overlap_read(v1, read, live, grow)
                    return True
            elif i21 > -1 and i22 > -1:
                if (i11 == -1 and i12 == -1) or abs(i21 - i22) < abs(i11 - i12):
                    self._read_overlap_contig(v2, read, live, grow)
                    return True
            return False
        if s1 > s2:
            self._contig_overlap_read(v1, read, live, grow)
        else:
            self._read_overlap_contig(v2, read, live, grow)
        return True

    def _contig_overlap_read(self, aln, read, live, grow):            # :506-528
        if aln[2] == len(self.seq) and aln[3] == 0:
            self._set_superseq(read, aln[5], aln[4])
            if grow:
                self.set_kmers(live)
            return
        post = read.seq[aln[4]:]
        nseq = self.seq[len(self.seq) - (self.k - 1):] + post
        self.seq = self.seq + post                                     # add_postseq :243-250
        self._set_counts(aln[3], aln[2], read.nreads, read.indel_only)
        self._extend_counts(len(post), read.nreads, read.indel_only, True)
        if grow:
            self.kmers.extend(read_kmers_ordered(nseq, self.k, live, FOR))

    def _read_overlap_contig(self, aln, read, live, grow):            # :530-546
        if aln[2] == len(read.seq) and aln[3] == 0:
            self._set_counts(aln[5], aln[4], read.nreads, read.indel_only)
            return
        pre = read.seq[0:aln[3]]
        nseq = pre + self.seq[0:self.k - 1]
        self.seq = pre + self.seq                                      # add_preseq :255-262
        self._set_counts(aln[5], aln[4], read.nreads, read.indel_only)
        self._extend_counts(len(pre), read.nreads, read.indel_only, False)
        if grow:
            self.kmers.extend(read_kmers_ordered(nseq, self.k, live, REV))


class Assembler:
    """init_assembly (:30-63) and everything it drives."""

    def __init__(self, mers, reads, k, rc_thresh, read_len, nw=None, stats=None):
        self.k = k
        self.rc_thresh = int(rc_thresh)
        self.read_len = read_len
        self.reads = reads                         # list[UniqueRead], fq_recs order
        self.nw = nw or nw_py.nw_fast
        self.stats = stats if stats is not None else {}
        # kmers.add_kmer (:276-278, Q5) + get_all_kmer_values (:280-285, Q7)
        kept = [(int(c), m) for m, c in mers.items() if len(set(m)) > 1]
        kept.sort(reverse=True)
        self.alive = {m: c for c, m in kept}       # akmers.mers, ordered
        self.live = set()                          # akmers.smers_set
        self.used_mers = set()                     # buffer.used_mers
        self.queue = {}                            # buffer.contigs: read idx -> Contig (FIFO)
        self.n_mers_in = len(mers)

    def _bump(self, key, by=1):
        self.stats[key] = self.stats.get(key, 0) + by

    # ---- find_reads / read_search (:102-122, Q9, Q10, Q28) ------------------
    def find_reads(self, mer, exclude, rev=False):
        self._bump("find_reads")
        hits = []
        for r in self.reads:
            if r.deleted:
                continue
            p = r.seq.find(mer)
            if p >= 0 and r.idx not in exclude:
                hits.append((r, p))
        if rev:
            hits.sort(key=lambda h: (-h[1], -len(h[0].seq)))
        else:
            hits.sort(key=lambda h: (h[1], -len(h[0].seq)))
        return hits

    # ---- buffer (:331-366) ------------------------------------------------
    def _add_contig(self, read, ct):                                   # :337-340
        if read.idx not in self.queue and not read.used:
            self.queue[read.idx] = ct
            read.used = True

    # ---- contig.check_read (:552-566, Q30) ----------------------------------
    def _check_read(self, ct, mer, read, grow):
        ct.buffer.add(read.idx)
        self._bump("check_align")
        self._bump("cells", 2 * len(ct.seq) * len(read.seq))
        match = ct.check_align(read, mer, self.live, grow, self.nw)
        if match:
            read.used = True
            ct.batch.append((read, True))
        else:
            ct.batch.append((read, False))
            if self.alive[mer] > 2 and not read.used:
                ct.alt.append(read)
            else:
                ct.delete.append(read)
        return match

    # ---- contig.check_alt_reads (:568-582, Q12, Q13) ------------------------
    def _check_alt_reads(self, ct):
        new = []
        taken = set()
        for read in ct.alt:
            x = read_kmer_set(read.seq, self.k, self.live) - self.used_mers - taken
            if x:
                for mer in sorted(x):                                  # order policy Q13
                    if self.alive[mer] > 1:
                        new.append((read, Contig(mer, read, self.k)))
                        taken |= x
                        break
        return new

    # ---- contig.finalize (:584-599) + read_batch.clean (:389-397, Q11) -------
    def _finalize(self, ct, setup):
        if setup:
            ct.set_kmers(self.live)
        for read, nc in self._check_alt_reads(ct):
            self._add_contig(read, nc)
        keep = [b for b in ct.batch if b[1]]
        ct.reads |= {b[0].idx for b in keep}
        for read in ct.delete:
            read.deleted = True                    # del fq_recs[read.seq]
        ct.delete = []
        ct.alt = []
        ct.batch = [keep[-1]]

    # ---- setup_contigs (:11-26, Q20) ------------------------------------------
    def _setup_contigs(self, mer):
        ct = None
        hits = self.find_reads(mer, ())
        self.used_mers.add(mer)
        for read, _pos in hits:
            if ct is None:
                ct = Contig(mer, read, self.k)
                self._add_contig(read, ct)
            else:
                self._check_read(ct, mer, read, False)
        if ct is not None:
            self._finalize(ct, True)

    # ---- contig.grow (:616-649) ---------------------------------------------
    def _grow(self, ct):
        if not ct.setup:
            ct.set_kmers(self.live)
        while True:
            done = set(ct.checked)
            todo = [t for t in ct.kmers if t[0] not in done]           # refresh_kmers :601
            if not todo:
                break
            for mer, _pos, lth, _dist, order in todo:
                if order == MID:                                       # get_mer_reads :604-614
                    rev = (lth == 0)
                else:
                    rev = (order == FOR)
                hits = self.find_reads(mer, ct.buffer, rev)
                self.used_mers.add(mer)
                for read, _p in hits:
                    if self._check_read(ct, mer, read, True):
                        self.queue.pop(read.idx, None)                 # buff.remove_contig :639
                self._finalize(ct, False)
                ct.checked.append(mer)
        ct.set_kmer_locs()

    def run(self):
        out = []
        if self.n_mers_in == 0:                                        # :33-34
            return out
        while self.alive and max(self.alive.values()) > 1:             # has_mers :318-322
            self.live = set(self.alive)                                # update_smer_set
            mer = next(iter(self.alive))
            self._bump("seeds")
            self._setup_contigs(mer)
            while self.queue:
                ridx = next(iter(self.queue))                          # get_contig :346-350
                ct = self.queue.pop(ridx)
                self._grow(ct)
                if ct.total_reads() < self.rc_thresh or len(ct.seq) <= self.read_len:   # :53 (Q21)
                    continue
                out.append(ct)
            for m in self.used_mers:                                   # remove_kmers :358-360
                del self.alive[m]
            self.used_mers = set()
        return out


def contig_record(ct, reads):
    """Canonical, comparable form of one contig (what sv_processor.contig reads,
    SURVEY.md section 3.5).  `reads` compared as a sorted id list (Q22)."""
    ids = sorted(rid for u in ct.reads for rid in [reads[u].rep_id])
    return {
        "seq": ct.seq,
        "indel_only": list(ct.indel_only),
        "others": list(ct.others),
        "reads": ids,
        "kmers": [[t[0], t[1], t[2], t[3], t[4]] for t in ct.kmers],
        "kmer_locs": list(ct.kmer_locs),
    }


def init_assembly(mers, records, k, rc_thresh, read_len, nw=None, stats=None):
    """records: (id, seq, qual, indel_only) in cleaned-FASTQ order.
    Returns the list of contig records in acceptance order."""
    reads = group_reads(records)
    asm = Assembler(mers, reads, k, rc_thresh, read_len, nw=nw, stats=stats)
    return [contig_record(ct, reads) for ct in asm.run()]
