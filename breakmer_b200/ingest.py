"""Native ingest: the FASTA / FASTQ texts of a batch
of targets -> the packed arrays of bk_compare_kmers_batch, parsed by the library's host
threads straight into page-locked memory (bk_ingest_buffers / bk_ingest_files,
include/breakmer_b200.h).  Stands in for FastqFile (utils.py:692-720), the record model
of get_fastq_reads (utils.py:230-244) and the readers inside `jellyfish count`
(utils.py:160) on the way into target.compare_kmers.

    ing = Ingest()                                   # one per host thread that submits batches
    pk = ing.files(refs, reads, scs, k=15, rc_thresh=2)   # or ing.texts(...) for in-memory text
    out = batch.run(handle, pk)

`IngestedBatch` has the interface of batch.PackedBatch (struct(), n, k, read_ids, ...); its
arrays live in the Ingest object's buffer and are valid until the next call on it.
"""
import ctypes
from ctypes import byref, c_char_p, c_void_p

import numpy as np

from . import _lib


class IngestedBatch:
    def __init__(self, owner, s, text, n, k, rc_thresh, names, has_normal, with_ref):
        self._owner = owner            # keeps the buffer alive
        self._s = s
        self._text = text
        self.n = n
        self.k = int(k)
        self.rc_thresh = int(rc_thresh)
        self.names = names if names is not None else [str(i) for i in range(n)]
        self.has_normal = has_normal
        self.with_ref = with_ref
        self.in_mers = None
        self.n_reads = int(text.n_reads)
        self._ids = None
        s.k = self.k
        s.rc_thresh = self.rc_thresh
        if not with_ref:
            s.ref_bases = None
            s.ref_off = None
        self.read_reg_off = self._view(s.read_reg_off, n + 1, np.int64)
        ro = self._view(s.read_off, self.n_reads + 1, np.int64)
        self.input_bytes = int(ro[-1]) if self.n_reads else 0          # bases of all inputs, like PackedBatch.input_bytes
        if with_ref and n:
            self.input_bytes += int(self._view(s.ref_off, n + 1, np.int64)[-1])
        for offp, regp in ((s.sc_off, s.sc_reg_off), (s.normal_off, s.normal_reg_off) if has_normal else (None, None)):
            if offp and regp and n:
                n_rec = int(self._view(regp, n + 1, np.int64)[-1])
                self.input_bytes += int(self._view(offp, n_rec + 1, np.int64)[-1])
        self.read_len = self._view(s.read_len, n + 1, np.int32)
        # mutable: the caller may replace the flags parsed from the "_1" header suffix (utils.py:436-443) with its own
        # fq_read.indel_only values (get_fastq_reads takes them from sv_reads, utils.py:215,240)
        self.read_flags = self._view(ctypes.cast(text.read_flags, c_void_p).value, max(self.n_reads, 1), np.uint8)

    @staticmethod
    def _view(addr, count, dtype):
        if isinstance(addr, ctypes._Pointer):
            addr = ctypes.cast(addr, c_void_p).value
        if not addr or count <= 0:
            return np.zeros(0, dtype)
        buf = (ctypes.c_uint8 * (count * np.dtype(dtype).itemsize)).from_address(addr)
        return np.frombuffer(buf, dtype=dtype, count=count)

    def struct(self):
        return self._s

    def _strings(self, bytes_ptr, off_ptr):
        off = self._view(off_ptr, self.n_reads + 1, np.int64)
        if self.n_reads == 0:
            return []
        raw = self._view(bytes_ptr, int(off[-1]), np.uint8).tobytes()
        return [raw[off[i]:off[i + 1]].decode() for i in range(self.n_reads)]

    @property
    def read_ids(self):
        """Record headers (fq_read.id), decoded on first use."""
        if self._ids is None:
            self._ids = self._strings(self._text.id_bytes, self._text.id_off)
        return self._ids

    def read_quals(self):
        return self._strings(self._text.qual_bytes, self._text.qual_off)

    def read_seqs(self):
        off = self._view(self._s.read_off, self.n_reads + 1, np.int64)
        if self.n_reads == 0:
            return []
        raw = self._view(self._s.read_bases, int(off[-1]), np.uint8).tobytes()
        return [raw[off[i]:off[i + 1]].decode() for i in range(self.n_reads)]

    def sequences(self, which):
        """Records of one k-mer input ('ref' | 'sc' | 'normal') per region, for tests."""
        s = self._s
        if which == "ref":
            off = self._view(s.ref_off, self.n + 1, np.int64)
            raw = self._view(s.ref_bases, int(off[-1]), np.uint8).tobytes()
            return [[raw[off[r]:off[r + 1]].decode()] for r in range(self.n)]
        bases, offp, regp = ((s.sc_bases, s.sc_off, s.sc_reg_off) if which == "sc" else
                             (s.normal_bases, s.normal_off, s.normal_reg_off))
        reg = self._view(regp, self.n + 1, np.int64)
        if len(reg) == 0:
            return [[] for _ in range(self.n)]
        off = self._view(offp, int(reg[-1]) + 1, np.int64)
        raw = self._view(bases, int(off[-1]), np.uint8).tobytes()
        return [[raw[off[i]:off[i + 1]].decode() for i in range(int(reg[r]), int(reg[r + 1]))] for r in range(self.n)]


class Ingest:
    """bk_ingest_create / bk_ingest_destroy.  pinned=True needs a CUDA device (page-locked buffer);
    pinned=False is plain host memory (parsing itself never touches the device)."""

    def __init__(self, n_threads=0, pinned=True):
        self.lib = _lib.load()
        self.g = c_void_p()
        rc = self.lib.bk_ingest_create(int(n_threads), 1 if pinned else 0, byref(self.g))
        if rc != _lib.BK_OK:
            raise _lib.BreakmerError(rc, "bk_ingest_create failed (pinned buffers need a CUDA device)")

    def close(self):
        if self.g:
            self.lib.bk_ingest_destroy(self.g)
            self.g = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != _lib.BK_OK:
            msg = self.lib.bk_ingest_last_error(self.g).decode()
            if rc == _lib.BK_ERR_FORMAT:
                raise ValueError(msg)                     # what FastqFile.next raises (utils.py:704-719)
            if rc == _lib.BK_ERR_IO:
                raise IOError(msg)
            raise _lib.BreakmerError(rc, msg)

    def texts(self, ref, reads, sc, normal=None, k=15, rc_thresh=2, names=None, with_ref=True):
        """Each argument: list of bytes/str (or None entries), one per region."""
        n = len(reads)
        keep = []

        def arr(lst):
            if lst is None:
                return None
            a = (_lib.Text * max(n, 1))()
            for i, t in enumerate(lst):
                if t is None:
                    continue
                b = t.encode() if isinstance(t, str) else bytes(t)
                keep.append(b)
                a[i].p = ctypes.cast(ctypes.c_char_p(b), c_void_p)
                a[i].n = len(b)
            return a

        s = _lib.BatchInput()
        text = _lib.IngestText()
        self._check(self.lib.bk_ingest_buffers(self.g, n, arr(ref), arr(reads), arr(sc), arr(normal), byref(s), byref(text)))
        return IngestedBatch(self, s, text, n, k, rc_thresh, names, normal is not None, with_ref)

    def files(self, ref, reads, sc, normal=None, k=15, rc_thresh=2, names=None, with_ref=True):
        """Each argument: list of paths (None / "" = absent), one per region."""
        n = len(reads)

        def arr(lst):
            if lst is None:
                return None
            a = (c_char_p * max(n, 1))()
            for i, p in enumerate(lst):
                a[i] = p.encode() if p else None
            return a

        s = _lib.BatchInput()
        text = _lib.IngestText()
        self._check(self.lib.bk_ingest_files(self.g, n, arr(ref), arr(reads), arr(sc), arr(normal), byref(s), byref(text)))
        return IngestedBatch(self, s, text, n, k, rc_thresh, names, normal is not None, with_ref)

    def write_contigs(self, res, pk, contigs_dirs, cluster_fns=None):
        """bk_write_contigs: the contig.setup files (sv_processor.py:749-782) of every contig of `res`, the raw
        result (batch.run(..., decode=False)) of the call made with the IngestedBatch `pk`.  Returns the number
        of files written."""
        n = pk.n

        def arr(lst):
            if lst is None:
                return None
            a = (c_char_p * max(n, 1))()
            for i, p in enumerate(lst):
                a[i] = p.encode() if p else None
            return a

        nf = ctypes.c_int64(0)
        s = pk.struct()
        self._check(self.lib.bk_write_contigs(self.g, byref(res), byref(s), byref(pk._text), arr(contigs_dirs),
                                              arr(cluster_fns), byref(nf)))
        return int(nf.value)

    def write_sample_kmers(self, res, k, paths):
        """bk_write_sample_kmers: "<mer>\\t<count>" files of every target (sv_processor.py:625-632) from the raw result."""
        n = int(res.n_regions)
        a = (c_char_p * max(n, 1))()
        for i, p in enumerate(paths):
            a[i] = p.encode() if p else None
        nf = ctypes.c_int64(0)
        self._check(self.lib.bk_write_sample_kmers(self.g, byref(res), int(k), a, byref(nf)))
        return int(nf.value)
