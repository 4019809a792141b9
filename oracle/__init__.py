"""CPU oracle for the BreaKmer per-target k-mer assembly hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``breakmer_b200`` may import, link or
execute anything in this package.  The only legitimate users are ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` -- and there only as the checker / reported baseline,
never as the thing shipped.

Parity pinning: the reference (Python 2, ``/root/reference``) has no tests or
golden vectors of its own.  ``oracle/ref_shim.py`` loads the reference's own
``olc.py`` and ``sv_assembly.py`` *from where they lie* under
``/root/reference`` with a minimal Python-3 shim applied in memory, and
``oracle/make_golden.py`` uses that to generate the fixtures committed under
``tests/golden/``.  The restatements in this package are checked against
those fixtures (``tests/test_oracle_*.py``).  ``ref_shim.load_mm2`` does the same for
``sv_assembly_mm2.py`` (read-redundancy row, ``redundancy_py.py``).  K-mer counting follows
jellyfish 1.1.11 (third party, absent from the reference tree): that part is
"parity unpinned" -- see ``kmers_py.py``.
"""
