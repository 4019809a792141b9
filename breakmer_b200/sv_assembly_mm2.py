"""Read-redundancy operators of the reference's older assembler variant, on the GPU
(SURVEY.md section 8.7 f.4).

Mirrors the names of /root/reference/sv_assembly_mm2.py (and of the dead copies in
sv_assembly.py:67-98):

    same_reads(seq1, seq2)            :64-70    bool
    subseq(seq1, seq2)                :74-85    (bool, None | score); threshold SUBSEQ_FRAC
    sim_seqs(seq1, b_read)            :88-94    bool
    b_read, read_batch                :281-355  read_batch.check_mer_read(pos, read) one read at a time
                                                (called from contig.check_read, :478; batch opened at :368)

and adds the batched entry the GPU wants:

    dedup_batches(batches)            every batch's check_mer_read chain from ONE alignment launch
                                      (bk_dedup_reads: all ordered pairs of a batch through the olc.nw
                                      kernel, decision chain replayed on the host in C++)

Only the de-duplication step is built; the rest of the mm2 assembler (stricter acceptance
thresholds, single counts vector) is imported by nothing in the reference and is not rebuilt.

Kept quirk: ``subseq`` returns a tuple and ``sim_seqs`` tests ``same_reads(..) or subseq(..)``, so
``sim_seqs`` is True for every b_read that is not flagged redundant (sv_assembly_mm2.py:92).
"""
import numpy as np

from . import _lib, get_handle, olc

SUBSEQ_FRAC = 0.90            # sv_assembly_mm2.py:77 (sv_assembly.py:80 uses 0.85)

ADDED, REDUNDANT, DELETED = 1, 2, 4     # BK_DEDUP_* of include/breakmer_b200.h


def same_reads_batch(pairs, device=0):
    res = olc.nw_batch(pairs, want_aln=False, device=device)
    return [aln[3] == 0 and aln[5] == 0 and aln[6] > 0.95 * len(a) for aln, (a, _b) in zip(res, pairs)]


def subseq_batch(pairs, frac=SUBSEQ_FRAC, device=0):
    """[(seq1, seq2), ...] -> [subseq(seq1, seq2), ...]: is seq2 contained in seq1?"""
    res = olc.nw_batch([(b, a) for a, b in pairs], want_aln=False, device=device)
    out = []
    for aln, (seq1, seq2) in zip(res, pairs):
        if aln[2] == len(seq2) and aln[3] == 0 and aln[6] >= frac * len(seq2):
            out.append((True, None) if len(seq2) < len(seq1) else (True, aln[6]))
        else:
            out.append((False, aln[6]))
    return out


def same_reads(seq1, seq2):
    return same_reads_batch([(seq1, seq2)])[0]


def subseq(seq1, seq2, frac=SUBSEQ_FRAC):
    return subseq_batch([(seq1, seq2)], frac)[0]


class b_read:
    def __init__(self, read, redundant, checked, aligned):
        self.read = read
        self.redundant = redundant
        self.align_checked = checked
        self.aligned = aligned


def sim_seqs(seq1, b_read, frac=SUBSEQ_FRAC):
    if b_read.redundant:
        return False
    seq2 = b_read.read.seq
    return bool(same_reads(seq1, seq2) or subseq(seq2, seq1, frac))


class read_batch:
    """Incremental form with the reference's interface: one small launch per decision.  Use
    dedup_batches when the reads of a batch are known up front."""

    def __init__(self, read, mer_pos):
        self.delete = set()
        self.alt = []
        self.batch_reads = [b_read(read, False, True, True)]
        self.mer_pos_d = {mer_pos: [0]}

    def set_last_read_aligned(self):
        self.batch_reads[-1].aligned = True

    def check_mer_read(self, pos, read, frac=SUBSEQ_FRAC):
        if pos in self.mer_pos_d and any(not self.batch_reads[x].redundant for x in self.mer_pos_d[pos]):
            self.delete.add(read.id)                   # sim_seqs is True for each of those (module docstring)
            return False
        last = self.batch_reads[-1]
        ss1, ss2 = subseq_batch([(last.read.seq, read.seq), (read.seq, last.read.seq)], frac)
        if ss1[0] and not ss1[1]:
            self.delete.add(read.id)
            return False
        if ss2[0] and not ss2[1]:
            self.delete.add(last.read.id)
            last.redundant = True
        elif (ss1[0] and ss1[1]) or (ss2[0] and ss2[1]):
            if ss1[0] and ss1[1] >= ss2[1]:
                self.delete.add(read.id)
                return False
            if ss2[0] and ss2[1] >= ss1[1]:
                self.delete.add(last.read.id)
                last.redundant = True
        self.mer_pos_d.setdefault(pos, []).append(len(self.batch_reads))
        self.batch_reads.append(b_read(read, False, True, False))
        return True


def dedup_batches(batches, frac=SUBSEQ_FRAC, device=0):
    """batches: [[(read, mer_pos), ...], ...] -- for each seed k-mer the reads in the order the
    assembler meets them (the first opens the batch); `read` has .id and .seq.
    -> per batch a dict: checks (check_mer_read's return per read, True for the opener),
    kept (reads appended to batch_reads, in order), redundant (kept reads flagged redundant later),
    deleted (ids in read_batch.delete)."""
    if not batches:
        return []
    seqs, pos, off = [], [], [0]
    for b in batches:
        if not b:
            raise ValueError("dedup_batches: empty batch")
        for read, p in b:
            seqs.append(read.seq)
            pos.append(int(p))
        off.append(len(seqs))
    try:
        check, flags, _n_pairs = get_handle(device).dedup_reads(seqs, pos, off, frac)
    except _lib.BreakmerError as e:
        if e.code == _lib.BK_ERR_EMPTY_SEQ:
            raise NameError(str(e))
        raise
    out = []
    for bi, b in enumerate(batches):
        lo = off[bi]
        fl = flags[lo:lo + len(b)]
        out.append({
            "checks": [bool(c) for c in check[lo:lo + len(b)]],
            "kept": [b[i][0] for i in np.flatnonzero(fl & ADDED)],
            "redundant": [b[i][0] for i in np.flatnonzero(fl & REDUNDANT)],
            "deleted": {b[i][0].id for i in np.flatnonzero(fl & DELETED)},
        })
    return out
