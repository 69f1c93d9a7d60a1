#!/usr/bin/env python3
"""Records what the REFERENCE'S OWN unit tests do to class seqComparer, as data, so that the same call sequences can be
replayed against the CUDA engine on the GPU box (where /root/reference does not exist).

    python tests/golden/make_ref_trace.py        # build container only; writes tests/golden/ref_unittest_trace.json

The unmodified reference module (/root/reference/src/seqComparer.py, through oracle/ref_shim.py) runs its 65
unittest.TestCase classes (:1130-2257) with `seqComparer` replaced by a recording proxy around the real class: every
construction, method call (arguments, return value or exception type), attribute assignment (`sc._seq1 = ...`) and read of
`seqProfile[...]` / `consensi.keys()` becomes one event.  tests/test_ref_trace_gpu.py replays the events against
findneighbour3_b200.seqComparer and compares every return value.  No reference source is copied: the file holds the
tests' INPUTS and the reference's OUTPUTS.
"""
import json
import os
import sys
import types
import unittest

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402

# results that depend on random.sample over a set of str guids (hash-seed dependent): replayed, compared in structure and in
# their deterministic fields only (tests/golden/msa.json and cluster.json pin the numbers with fixed inputs)
NONDET = {"estimate_expected_unk", "estimate_expected_unk_sites", "multi_sequence_alignment", "_msa"}


def enc(v):
    """value -> JSON-friendly, type-preserving"""
    if v is None or isinstance(v, (bool, str)):
        return v
    if isinstance(v, (int, np.integer)):
        return int(v)
    if isinstance(v, (float, np.floating)):
        return {"__float__": repr(float(v))}
    if isinstance(v, set):
        return {"__set__": sorted((enc(x) for x in v), key=repr)}
    if isinstance(v, frozenset):
        return {"__set__": sorted((enc(x) for x in v), key=repr)}
    if isinstance(v, tuple):
        return {"__tuple__": [enc(x) for x in v]}
    if isinstance(v, list):
        return [enc(x) for x in v]
    if isinstance(v, dict):
        return {"__dict__": [[enc(k), enc(x)] for k, x in sorted(v.items(), key=lambda kv: repr(kv[0]))]}
    if type(v).__name__ == "DataFrame":
        return {"__df__": enc(v.to_dict(orient="index"))}
    if type(v).__name__ == "Seq":
        return str(v)
    return {"__repr__": repr(v)[:200]}


class DictProxy:
    def __init__(self, rec, attr, real):
        self._rec, self._attr, self._real = rec, attr, real

    def __getitem__(self, k):
        v = self._real[k]
        self._rec.event({"op": "getitem", "attr": self._attr, "key": enc(k), "val": enc(v)})
        return v

    def __contains__(self, k):
        r = k in self._real
        self._rec.event({"op": "contains", "attr": self._attr, "key": enc(k), "val": r})
        return r

    def __len__(self):
        r = len(self._real)
        self._rec.event({"op": "len", "attr": self._attr, "val": r})
        return r

    def keys(self):
        r = self._real.keys()
        self._rec.event({"op": "keys", "attr": self._attr, "val": enc(set(r))})
        return r

    def __iter__(self):
        return iter(self._real)


class Trace:
    def __init__(self):
        self.events = []
        self.n_obj = 0


def make_recorder(real_class, trace):
    class Recorder:
        def __init__(self, *a, **kw):
            object.__setattr__(self, "_id", trace.n_obj)
            trace.n_obj += 1
            ev = {"op": "new", "id": self._id, "args": enc(list(a)), "kwargs": enc(kw)}
            try:
                object.__setattr__(self, "_obj", real_class(*a, **kw))
            except Exception as ex:
                ev["exc"] = type(ex).__name__
                trace.events.append(ev)
                raise
            trace.events.append(ev)

        def event(self, ev):
            ev["id"] = self._id
            trace.events.append(ev)

        def __getattr__(self, name):
            val = getattr(self._obj, name)
            if isinstance(val, types.MethodType):
                def call(*a, **kw):
                    ev = {"op": "call", "name": name, "args": enc(list(a)), "kwargs": enc(kw)}
                    try:
                        r = val(*a, **kw)
                    except Exception as ex:
                        ev["exc"] = type(ex).__name__
                        self.event(ev)
                        raise
                    if isinstance(r, types.GeneratorType):
                        r = list(r)
                        ev["ret"] = {"__gen__": enc(r)}
                        self.event(ev)
                        return iter(r)
                    ev["ret"] = enc(r)
                    if name in NONDET:
                        ev["nondet"] = True
                    self.event(ev)
                    return r
                return call
            if name in ("seqProfile", "consensi"):
                return DictProxy(self, name, val)
            self.event({"op": "get", "name": name, "val": enc(val)})
            return val

        def __setattr__(self, name, value):
            self.event({"op": "set", "name": name, "val": enc(value)})
            setattr(self._obj, name, value)

    return Recorder


def main():
    ref = ref_shim.load_reference_module("seqComparer")
    real = ref.seqComparer
    out = []
    for name in sorted(dir(ref)):
        obj = getattr(ref, name)
        if not (isinstance(obj, type) and issubclass(obj, unittest.TestCase) and name.startswith("test_")):
            continue
        trace = Trace()
        ref.seqComparer = make_recorder(real, trace)
        try:
            res = unittest.TestResult()
            unittest.defaultTestLoader.loadTestsFromTestCase(obj).run(res)
        finally:
            ref.seqComparer = real
        if res.errors or res.failures:
            print("skipped (fails against the reference itself here):", name,
                  (res.errors + res.failures)[0][1].strip().splitlines()[-1][:100])
            continue
        out.append({"test": name, "events": trace.events})
    path = os.path.join(HERE, "ref_unittest_trace.json")
    with open(path, "w") as fh:
        json.dump({"source": "src/seqComparer.py:1130-2257 (unittest classes of the unmodified reference)", "tests": out}, fh,
                  separators=(",", ":"))
    n_ev = sum(len(t["events"]) for t in out)
    print("wrote %s: %d tests, %d events, %d bytes" % (path, len(out), n_ev, os.path.getsize(path)))


if __name__ == "__main__":
    main()
