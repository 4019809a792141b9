"""Load the *unmodified* reference modules from /root/reference under Python 3.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by
``oracle/make_golden.py`` in the build container, where ``/root/reference``
exists, to produce the golden fixtures under ``tests/golden/``.  It is never
used at run time on the GPU box (the reference tree does not exist there).

The reference is Python 2.  Its two hot-path modules,
``olc.py`` and ``sv_assembly.py``, are read from the reference tree, a short
list of *syntactic* Python-2 -> Python-3 rewrites is applied to the text in
memory, and the result is exec'd.  No reference source is written to disk or
copied into this repository.  Rewrites (each is behaviour-preserving with
respect to CPython 2.7):

  olc.py
    * expandtabs(8)          -- py2 treats a tab as 8 columns (olc.py:66-74)
  sv_assembly.py
    * drop the ``__main__`` driver (sv_assembly.py:664-691; it holds a py2
      print statement and is broken anyway)
    * ``from utils import *``     -> removed; ``fq_read`` is supplied below
      (restating utils.py:681-688, the only name the live path needs)
    * ``lambda (x,y): x+y``       -> ``lambda x_y: x_y[0]+x_y[1]``  (:171,190,191)
    * ``len(seq)/2``              -> ``len(seq)//2``                 (:129)
    * ``.items()[0]``/``.keys()[0]`` -> ``list(...)[0]``             (:45,:347)
    * builtins ``map``/``filter``/``zip`` are shadowed in the module by eager,
      list-returning versions (py2 semantics; :359,:364,:390 rely on the side
      effects)

Two behaviours of the reference depend on CPython-2 hash iteration order and
cannot be observed anywhere.  The order policy of this
repository is applied to the reference in the same way:

  * Q9  -- ``fq_recs.items()`` order is insertion order (what dict gives on
           Python >= 3.7); nothing to rewrite.
  * Q13 -- ``for mer in list(x)`` in ``check_alt_reads`` (:575) iterates a set;
           rewritten to ``sorted(x)`` (smallest mer first).
"""
import builtins
import os
import re
import sys
import types

REFERENCE_ROOT = os.environ.get("BREAKMER_REFERENCE_ROOT", "/root/reference")


class fq_read:
    """Restates utils.py:681-688 (the record type the assembler consumes)."""

    def __init__(self, header, seq, qual, indel_only):
        self.id = header
        self.seq = str(seq)
        self.qual = str(qual)
        self.used = False
        self.dup = False
        self.indel_only = indel_only


def _eager_map(*a):
    return list(builtins.map(*a))


def _eager_filter(f, it):
    return list(builtins.filter(f, it))


def _eager_zip(*a):
    return list(builtins.zip(*a))


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sv_assembly.py"))


def _must_sub(pattern, repl, text, count_expected, flags=0):
    new, n = re.subn(pattern, repl, text, flags=flags)
    if n != count_expected:
        raise RuntimeError(
            "ref_shim: rewrite %r matched %d times, expected %d" % (pattern, n, count_expected))
    return new


def load():
    """Return (olc_module, sv_assembly_module) built from the reference tree."""
    if not available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)

    with open(os.path.join(REFERENCE_ROOT, "olc.py")) as f:
        olc_src = f.read().expandtabs(8)
    olc = types.ModuleType("olc")
    olc.__file__ = os.path.join(REFERENCE_ROOT, "olc.py")
    exec(compile(olc_src, olc.__file__, "exec"), olc.__dict__)

    with open(os.path.join(REFERENCE_ROOT, "sv_assembly.py")) as f:
        src = f.read().expandtabs(8)
    cut = src.index("if __name__ == '__main__'")
    src = src[:cut]
    src = _must_sub(r"^from utils import \*\s*$", "", src, 1, flags=re.M)
    src = _must_sub(r"lambda \(x,y\): x\+y", "lambda x_y: x_y[0]+x_y[1]", src, 3)
    src = _must_sub(r"m = len\(seq\)/2", "m = len(seq)//2", src, 1)
    src = _must_sub(r"akmers\.mers\.items\(\)\[0\]", "list(akmers.mers.items())[0]", src, 1)
    src = _must_sub(r"self\.contigs\.keys\(\)\[0\]", "list(self.contigs.keys())[0]", src, 1)
    src = _must_sub(r"for mer in list\(x\) :", "for mer in sorted(x) :", src, 1)

    asm = types.ModuleType("sv_assembly")
    asm.__file__ = os.path.join(REFERENCE_ROOT, "sv_assembly.py")
    asm.__dict__.update(map=_eager_map, filter=_eager_filter, zip=_eager_zip, fq_read=fq_read)
    saved = sys.modules.get("olc")
    sys.modules["olc"] = olc
    try:
        exec(compile(src, asm.__file__, "exec"), asm.__dict__)
    finally:
        if saved is None:
            del sys.modules["olc"]
        else:
            sys.modules["olc"] = saved
    return olc, asm


def load_mm2():
    """Return (olc_module, sv_assembly_mm2_module): the reference's older assembler variant
    (sv_assembly_mm2.py, imported by nothing in the reference), loaded the same way as ``load``.
    Used to pin the read-redundancy row: ``same_reads``, ``subseq``,
    ``sim_seqs`` (:64-93) and ``read_batch.check_mer_read`` (:309-355).  Rewrites: the ``__main__``
    driver (:577-) and ``from utils import *`` are dropped, ``len(seq)/2`` -> ``//`` (:125),
    ``.items()[0]``/``.keys()[0]`` -> ``list(...)[0]`` (:45,:257); ``map``/``filter``/``zip`` are eager."""
    fn = os.path.join(REFERENCE_ROOT, "sv_assembly_mm2.py")
    if not os.path.isfile(fn):
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    olc, _asm = load()
    with open(fn) as f:
        src = f.read().expandtabs(8)
    src = src[:src.index("if __name__ == '__main__'")]
    src = _must_sub(r"^from utils import \*\s*$", "", src, 1, flags=re.M)
    src = _must_sub(r"m = len\(seq\)/2", "m = len(seq)//2", src, 1)
    src = _must_sub(r"akmers\.mers\.items\(\)\[0\]", "list(akmers.mers.items())[0]", src, 1)
    src = _must_sub(r"self\.contigs\.keys\(\)\[0\]", "list(self.contigs.keys())[0]", src, 1)
    mm2 = types.ModuleType("sv_assembly_mm2")
    mm2.__file__ = fn
    mm2.__dict__.update(map=_eager_map, filter=_eager_filter, zip=_eager_zip, fq_read=fq_read)
    saved = sys.modules.get("olc")
    sys.modules["olc"] = olc
    try:
        exec(compile(src, fn, "exec"), mm2.__dict__)
    finally:
        if saved is None:
            del sys.modules["olc"]
        else:
            sys.modules["olc"] = saved
    return olc, mm2


def load_readers():
    """Namespace holding the reference's own ``fq_read``, ``FastqFile`` and
    ``get_fastq_reads`` (utils.py:203-246, 681-720), used to pin the ingest row
    (SURVEY.md section 8.7 f.1).  ``utils.py`` as a whole cannot be imported
    (pysam, Biopython), so the three definitions are cut out of the file's text by
    their delimiting comment rulers and exec'd; rewrites:

      * ``self._f.next()``  -> ``next(self._f)``  and ``__next__ = next``  (:703)
        (a StopIteration escaping the list comprehension still ends the iteration,
        on Python 3 as on 2.7)
      * ``open`` is shadowed by ``open(..., newline="\\n")``: CPython 2.7's text mode
        on Linux ends lines at "\\n" only (no universal-newline translation)
    """
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, "utils.py")):
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    with open(os.path.join(REFERENCE_ROOT, "utils.py")) as f:
        src = f.read().expandtabs(8)

    def cut(start_pat, end_pat):
        a = re.search(start_pat, src, flags=re.M)
        if a is None:
            raise RuntimeError("ref_shim: %r not found" % start_pat)
        b = re.search(end_pat, src[a.start():], flags=re.M)
        if b is None:
            raise RuntimeError("ref_shim: %r not found" % end_pat)
        return src[a.start():a.start() + b.start()]

    text = cut(r"^def get_fastq_reads\(fn, sv_reads\)", r"^#-{10,}")
    text += "\n" + cut(r"^class fq_read", r"^#@{10,}")
    fastq = cut(r"^class FastqFile", r"^# End FastqFile class")
    fastq = _must_sub(r"self\._f\.next\(\)", "next(self._f)", fastq, 1)
    fastq = _must_sub(r"^  def next\(self\) :", "  def __next__(self) :", fastq, 1, flags=re.M)
    text += "\n" + fastq
    ns = {"open": lambda fn, mode="r": builtins.open(fn, mode, newline="\n")}
    exec(compile(text, os.path.join(REFERENCE_ROOT, "utils.py"), "exec"), ns)
    return types.SimpleNamespace(fq_read=ns["fq_read"], FastqFile=ns["FastqFile"], get_fastq_reads=ns["get_fastq_reads"])


def load_contig_writers():
    """A class holding the reference's own ``setup`` / ``write_cluster_file`` / ``write_read_fq`` /
    ``write_contig_fa`` methods of ``sv_processor.contig`` (sv_processor.py:749-782), cut out of the file's
    text (the module as a whole needs pysam) -- used to pin the contig hand-off row.
    No rewrites are needed."""
    fn = os.path.join(REFERENCE_ROOT, "sv_processor.py")
    if not os.path.isfile(fn):
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    with open(fn) as f:
        src = f.read().expandtabs(8)
    a = re.search(r"^  def setup\(self, cluster_fn\) :", src, flags=re.M)
    b = re.search(r"^  def has_result\(self\) :", src, flags=re.M)
    if a is None or b is None or b.start() < a.start():
        raise RuntimeError("ref_shim: contig writers not found")
    text = "class contig_writers:\n" + src[a.start():b.start()]
    import logging
    ns = {"os": os, "logging": logging}
    exec(compile(text, fn, "exec"), ns)
    return ns["contig_writers"]
