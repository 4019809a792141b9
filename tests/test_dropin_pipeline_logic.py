"""Host logic of sv_processor.compare_kmers_batch's chunked pass (no device: the pipeline, the parser and the result
application are replaced by recording fakes).  What is under test is the scheduling: every target applied exactly once,
never more chunks un-applied than there are handles, errors of either thread reach the caller without a hang."""
import threading
import time

import pytest

from breakmer_b200 import sv_processor


class _T:
    def __init__(self, i, n_reads):
        self.name = "t%03d" % i
        self.cleaned_read_recs = {j: None for j in range(n_reads)}


class _FakePipe:
    def __init__(self, inflight, log):
        self.handles = list(range(inflight))
        self._queue = []
        self.log = log
        self.closed = False

    def full(self):
        return len(self._queue) >= len(self.handles)

    def pending(self):
        return len(self._queue)

    def submit(self, pk, tag=None):
        assert not self.full(), "submitted onto a busy handle"
        with self.log["lock"]:
            self.log["submitted"] += 1
            un_applied = self.log["submitted"] - self.log["applied"]
            self.log["max_unapplied"] = max(self.log["max_unapplied"], un_applied)
            assert un_applied <= len(self.handles), "a handle was reused before its chunk was applied"
        self._queue.append((pk, tag))

    def pop(self, decode=True):
        pk, tag = self._queue.pop(0)
        time.sleep(0.002)                                  # the device
        return ("res", pk), pk, tag

    def close(self):
        self.closed = True


def _install(monkeypatch, inflight, fail_apply_at=None, fail_pack_at=None):
    log = {"lock": threading.Lock(), "submitted": 0, "applied": 0, "max_unapplied": 0, "order": [], "threads": set(),
           "packed": 0, "slots": []}
    pipe = _FakePipe(inflight, log)
    monkeypatch.setattr(sv_processor, "_get_pipe", lambda dev, n: pipe)
    monkeypatch.setattr(sv_processor, "_get_ingest", lambda slot=0: "ingest-%s" % slot)

    def pack(targets, ingest, slot=0):
        log["packed"] += 1
        log["slots"].append(slot)
        if fail_pack_at is not None and log["packed"] == fail_pack_at:
            raise ValueError("pack failed")
        time.sleep(0.001)
        return [t.name for t in targets], [15] * len(targets), None

    def apply(targets, pk, res, ks, objs, ingest, write_contigs, failed, ing=None):
        assert res == ("res", pk) and pk == [t.name for t in targets]
        if fail_apply_at is not None and log["applied"] + 1 == fail_apply_at:
            raise RuntimeError("apply failed")
        time.sleep(0.001)
        with log["lock"]:
            log["applied"] += 1
            log["order"].extend(t.name for t in targets)
            log["threads"].add(threading.get_ident())
        if any(t.name == "t007" for t in targets):
            failed.append("t007")

    monkeypatch.setattr(sv_processor, "_pack_targets", pack)
    monkeypatch.setattr(sv_processor, "_apply_chunk", apply)
    return log, pipe


@pytest.mark.parametrize("apply_thread", [False, True])
@pytest.mark.parametrize("n,max_targets,inflight", [(50, 7, 3), (50, 7, 1), (23, 5, 8), (9, 2, 2)])
def test_every_target_applied_once(monkeypatch, apply_thread, n, max_targets, inflight):
    log, pipe = _install(monkeypatch, inflight)
    targets = [_T(i, (i * 37) % 11) for i in range(n)]
    with pytest.raises(sv_processor.CapacityError) as e:      # the fake reports t007 as over the limit
        sv_processor.compare_kmers_batch(targets, ingest="native", max_targets=max_targets, inflight=inflight,
                                         apply_thread=apply_thread)
    assert e.value.targets == ["t007"]
    assert sorted(log["order"]) == sorted(t.name for t in targets)
    assert log["submitted"] == log["applied"] == -(-n // (-(-n // -(-n // max_targets))))
    assert 1 <= log["max_unapplied"] <= inflight
    assert (threading.get_ident() in log["threads"]) == (not apply_thread)
    # parse buffers rotate so that none is reused while its chunk is still in flight
    assert set(log["slots"]) <= set(range(inflight + (1 if apply_thread else 0)))
    assert not pipe.closed and pipe.pending() == 0


@pytest.mark.parametrize("apply_thread", [False, True])
def test_errors_of_either_side_reach_the_caller(monkeypatch, apply_thread):
    log, pipe = _install(monkeypatch, 3, fail_apply_at=2)
    monkeypatch.setattr(sv_processor, "_pipes", {})
    targets = [_T(i, 3) for i in range(40)]
    with pytest.raises(RuntimeError, match="apply failed"):
        sv_processor.compare_kmers_batch(targets, ingest="native", max_targets=5, apply_thread=apply_thread)
    assert pipe.closed                                        # a pipeline with batches in flight is not kept
    log, pipe = _install(monkeypatch, 3, fail_pack_at=4)
    with pytest.raises(ValueError, match="pack failed"):
        sv_processor.compare_kmers_batch(targets, ingest="native", max_targets=5, apply_thread=apply_thread)
    assert pipe.closed
