"""Model-based test of the host layer (hypothesis RuleBasedStateMachine; CPU, numpy engine double).

Random sequences of the store's mutations -- store / overwrite, batch_store, bulk_load, delete, update_metadata, shard
clear, clear, save + reload (also onto a different device count) -- run against ``wdbx_b200.VectorStore`` and against a
plain-Python model of what the reference's semantics (plus DESIGN.md's documented decisions) say the store must hold.
After every step: count, get, unfiltered search, post-filtered search (the reference's per-shard candidate set,
vector_store.py:323-345) and the opt-in device pre-filter must agree with the model.  This is the class of bug ADVICE r1
found by hand (bulk ids after a shard clear, non-canonical bulk ids, prefix collisions)."""
import asyncio
import shutil
import tempfile

import numpy as np
from hypothesis import HealthCheck, settings, strategies as st
from hypothesis.stateful import RuleBasedStateMachine, initialize, invariant, rule

import wdbx_b200
from tests.fake_engine import FakeEngine
from wdbx_b200.shard_map import shard_for_id

DIM = 5
S = 3
EXPLICIT = [f"e{i}" for i in range(14)]


def _vec(seed: int) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal(DIM).astype(np.float32)


class Flaky(FakeEngine):
    """numpy engine double whose next append / overwrite / tombstone / clear can be made to fail (fault injection)"""
    fail = {}      # op -> fail on the n-th call from now (1 = the next one)

    def _maybe_fail(self, op):
        if op in Flaky.fail:
            Flaky.fail[op] -= 1
            if Flaky.fail[op] <= 0:
                del Flaky.fail[op]
                raise RuntimeError(f"injected device failure in {op}")

    def append(self, segment, rows, gids=None):
        self._maybe_fail("append")
        return super().append(segment, rows, gids=gids)

    def overwrite(self, segment, row, vector):
        self._maybe_fail("overwrite")
        return super().overwrite(segment, row, vector)

    def tombstone(self, segment, row, dead=True):
        self._maybe_fail("tombstone")
        return super().tombstone(segment, row, dead)

    def clear(self, segment=-1):
        self._maybe_fail("clear")
        return super().clear(segment)


class StoreMachine(RuleBasedStateMachine):
    devices = None          # subclass: GPU_DEVICES of the single-process multi-device layout

    def _open(self):
        cfg = {"GPU_STRICT": True}
        if self.devices:
            cfg["GPU_DEVICES"] = self.devices
        Flaky.fail = {}
        return wdbx_b200.VectorStore(DIM, self.dir, num_shards=S, config=wdbx_b200.WDBXConfig(cfg),
                                     dist=wdbx_b200.DistContext(0, 1, 0), _engine_factory=Flaky)

    @initialize()
    def start(self):
        # a small engine maximum: searches that ask for more than 8 neighbours are served by PAGING (allow bitmaps over the
        # rows not returned yet, vector_store._search_paged) -- the full-ranking probes below walk several pages
        import wdbx_b200._lib as lib

        self._max_k, lib.MAX_K = lib.MAX_K, 8
        self.dir = tempfile.mkdtemp(prefix="wdbx_model_")
        self.store = self._open()
        self.model = {}        # id -> dict(vec, meta, shard, order)   live rows only
        self.gone = set()      # ids that existed once and do not any more
        self.used_prefixes = set()
        self.explicit_live = set()   # live ids that are explicit rows (not rows of a bulk_load)
        self.order = 0         # insertion counter = the engine's gid order (tie rule)
        self.prefixes = 0
        self.seed = 1000

    def teardown(self):
        import wdbx_b200._lib as lib

        lib.MAX_K = getattr(self, "_max_k", lib.MAX_K)
        if getattr(self, "store", None) is not None:
            self.store.close()
        shutil.rmtree(getattr(self, "dir", ""), ignore_errors=True)

    def _next_seed(self):
        self.seed += 1
        return self.seed

    def _put(self, vid, vec, meta, shard_if_new):
        if vid in self.model:
            self.model[vid]["vec"] = vec          # overwrite in place: row (and tie order) kept
            self.model[vid]["meta"] = meta
        else:
            self.model[vid] = {"vec": vec, "meta": meta, "shard": shard_if_new, "order": self.order}
            self.gone.discard(vid)
            self.explicit_live.add(vid)
        self.order += 1

    # ------------------------------------------------------------------ mutations
    @rule(i=st.integers(0, len(EXPLICIT) - 1), g=st.integers(0, 2), with_meta=st.booleans())
    def store_one(self, i, g, with_meta):
        vid, vec = EXPLICIT[i], _vec(self._next_seed())
        meta = {"g": g} if with_meta else None
        assert self.store.store(vid, vec.tolist(), meta) is True
        self._put(vid, vec, meta or {}, shard_for_id(vid, S))

    @rule(target_bulk=st.booleans(), data=st.data())
    def store_existing_bulk_id(self, target_bulk, data):
        bulk_ids = [v for v in self.model if v not in self.explicit_live]
        if not bulk_ids:
            return
        vid = data.draw(st.sampled_from(sorted(bulk_ids)))
        vec = _vec(self._next_seed())
        assert self.store.store(vid, vec.tolist(), {"g": 1}) is True
        self._put(vid, vec, {"g": 1}, None)

    @rule(data=st.data())
    def store_a_gone_id_again(self, data):
        """an id that was deleted / cleared (a former bulk id included) comes back as a fresh explicit row"""
        if not self.gone:
            return
        vid = data.draw(st.sampled_from(sorted(self.gone)))
        vec = _vec(self._next_seed())
        assert self.store.store(vid, vec.tolist(), {"g": 2}) is True
        self._put(vid, vec, {"g": 2}, shard_for_id(vid, S))

    @rule(ids=st.lists(st.integers(0, len(EXPLICIT) - 1), min_size=1, max_size=5, unique=True), g=st.integers(0, 2))
    def batch(self, ids, g):
        vecs = {EXPLICIT[i]: _vec(self._next_seed()) for i in ids}
        meta = {EXPLICIT[i]: {"g": g} for i in ids[::2]}
        assert self.store.batch_store({k: v.tolist() for k, v in vecs.items()}, meta) == len(ids)
        # rows are appended shard by shard, in dict order inside a shard: that is the insertion (tie) order
        by_shard = {}
        for vid in vecs:
            shard = self.model[vid]["shard"] if vid in self.model else shard_for_id(vid, S)
            by_shard.setdefault(shard, []).append(vid)
        for shard, vids in by_shard.items():
            for vid in vids:
                self._put(vid, vecs[vid], meta.get(vid, {}), shard)

    @rule(n=st.integers(1, 13), family=st.sampled_from([None, None, "q", "q1", "q12", "e"]))
    def bulk(self, n, family):
        """`family`: prefixes that can collide -- with each other ("q1" + "23" = "q12" + "3") or with explicit ids ("e")"""
        import pytest

        if family is None:
            prefix = f"p{self.prefixes}_"
            self.prefixes += 1
        else:
            prefix = family
        X = np.stack([_vec(self._next_seed()) for _ in range(n)])

        def digits_only_extension(a, b):
            longer, shorter = (a, b) if len(a) > len(b) else (b, a)
            return longer.startswith(shorter) and longer[len(shorter):].isdigit()

        clash = any(prefix == p or digits_only_extension(prefix, p) for p in self.used_prefixes)
        clash = clash or any(v.startswith(prefix) and v[len(prefix):].isdigit() and str(int(v[len(prefix):])) == v[len(prefix):]
                             and int(v[len(prefix):]) < n for v in self.model if v in self.explicit_live)
        if clash:
            with pytest.raises(ValueError):
                self.store.bulk_load(X, id_prefix=prefix)
            return
        self.used_prefixes.add(prefix)
        assert self.store.bulk_load(X, id_prefix=prefix) == n
        base = self.order
        for i in range(n):
            self.model[f"{prefix}{i}"] = {"vec": X[i], "meta": {}, "shard": i % S, "order": base + i}
            self.gone.discard(f"{prefix}{i}")
        self.order += n

    @rule(i=st.integers(0, len(EXPLICIT) - 1), shard=st.integers(0, S - 1))
    def index_add(self, i, shard):
        """the operator boundary used directly: a new id lands in THAT shard (not its hash shard); an existing id is
        overwritten where it lives (the reference's per-shard indices know nothing of each other, but an id maps to one
        row in our store: DESIGN.md decision 2)"""
        vid, vec = EXPLICIT[i], _vec(self._next_seed())
        assert self.store.indices[shard].add(vid, vec) is True
        meta = self.model[vid]["meta"] if vid in self.model else {}
        self._put(vid, vec, meta, shard)

    @rule(data=st.data(), shard=st.integers(0, S - 1))
    def index_remove(self, data, shard):
        vid = data.draw(st.sampled_from(sorted(self.model) + ["e_unknown"]))
        here = vid in self.model and self.model[vid]["shard"] == shard
        assert self.store.indices[shard].remove(vid) is here
        if here:   # (the row's metadata goes with it: an id that is gone has none, whichever level removed it)
            del self.model[vid]
            self.gone.add(vid)
            self.explicit_live.discard(vid)

    @rule(kind=st.sampled_from(["short", "long", "text", "empty", "nested"]), i=st.integers(0, len(EXPLICIT) - 1))
    def store_invalid(self, kind, i):
        bad = {"short": [1.0] * (DIM - 1), "long": [1.0] * (DIM + 2), "text": "abc", "empty": [], "nested": [[1.0] * DIM]}[kind]
        self.store.strict = False
        try:
            assert self.store.store(EXPLICIT[i], bad, {"g": 0}) is False
            good = _vec(self._next_seed())
            other = EXPLICIT[(i + 1) % len(EXPLICIT)]
            assert self.store.batch_store({EXPLICIT[i]: bad, other: good.tolist()}, {other: {"g": 0}}) == 1
        finally:
            self.store.strict = True
        shard = self.model[other]["shard"] if other in self.model else shard_for_id(other, S)
        self._put(other, good, {"g": 0}, shard)

    @rule(what=st.sampled_from(["store_new", "store_existing", "delete", "clear_shard", "batch"]), data=st.data())
    def device_failure(self, what, data):
        """the engine call inside a mutation fails once (reference convention: logged, False returned): nothing of the
        half-done mutation may remain -- no id, no metadata change, no count change -- and the store keeps working"""
        if self.devices or len(getattr(self.store, "devices", [0])) > 1:
            return      # (a device group reserves room first and cannot roll back a stripe: multi_engine.py)
        self.store.strict = False
        try:
            if what == "store_new":
                fresh = [v for v in EXPLICIT if v not in self.model]
                if fresh:
                    Flaky.fail = {"append": 1}
                    assert self.store.store(fresh[0], _vec(self._next_seed()).tolist(), {"g": 1}) is False
            elif what == "store_existing" and self.model:
                vid = data.draw(st.sampled_from(sorted(self.model)))
                Flaky.fail = {"overwrite": 1}
                assert self.store.store(vid, _vec(self._next_seed()).tolist(), {"g": 1, "x": 1}) is False
            elif what == "delete" and self.model:
                vid = data.draw(st.sampled_from(sorted(self.model)))
                Flaky.fail = {"tombstone": 1}
                assert self.store.delete(vid) is False
            elif what == "clear_shard":
                Flaky.fail = {"clear": 1}
                assert self.store.indices[data.draw(st.integers(0, S - 1))].clear() is False
            elif what == "batch":
                fresh = [v for v in EXPLICIT if v not in self.model][:3]
                if fresh:
                    Flaky.fail = {"append": 1}
                    n = self.store.batch_store({v: _vec(self._next_seed()).tolist() for v in fresh}, {v: {"g": 1} for v in fresh})
                    # the first shard's append failed; the other shards' vectors went in
                    stored = [v for v in fresh if self.store._locate(v) is not None]
                    assert n == len(stored) < len(fresh)
                    for v in stored:
                        self._put(v, np.asarray(self.store.get(v)[0], np.float32), {"g": 1}, shard_for_id(v, S))
        finally:
            Flaky.fail = {}
            self.store.strict = True

    @rule(n=st.integers(S, 9), fail_at=st.integers(1, S))
    def bulk_load_fails_midway(self, n, fail_at):
        """the device append of shard `fail_at - 1` fails: the shards before it hold their rows, the ids of the others do
        not resolve, the prefix stays taken, and the error reaches the caller (bulk_load is the additive API: it raises)"""
        import pytest

        if len(getattr(self.store, "devices", [0])) > 1:
            return
        prefix = f"p{self.prefixes}_"
        self.prefixes += 1
        X = np.stack([_vec(self._next_seed()) for _ in range(n)])
        Flaky.fail = {"append": fail_at}
        try:
            with pytest.raises(RuntimeError, match="injected"):
                self.store.bulk_load(X, id_prefix=prefix)
        finally:
            Flaky.fail = {}
        self.used_prefixes.add(prefix)
        base = self.order
        for i in range(n):
            if i % S < fail_at - 1:
                self.model[f"{prefix}{i}"] = {"vec": X[i], "meta": {}, "shard": i % S, "order": base + i}
            else:
                self.gone.add(f"{prefix}{i}")
        self.order += n

    @rule(data=st.data())
    def delete(self, data):
        known = sorted(self.model) + ["e_unknown", "p0_999", "p0_01"]
        vid = data.draw(st.sampled_from(known))
        assert self.store.delete(vid) is (vid in self.model)
        if self.model.pop(vid, None) is not None:
            self.gone.add(vid)
            self.explicit_live.discard(vid)

    @rule(data=st.data(), g=st.integers(0, 2))
    def update_meta(self, data, g):
        known = sorted(self.model) + ["e_unknown"]
        vid = data.draw(st.sampled_from(known))
        assert self.store.update_metadata(vid, {"g": g, "u": True}) is (vid in self.model)
        if vid in self.model:
            self.model[vid]["meta"] = {"g": g, "u": True}

    @rule(shard=st.integers(0, S - 1))
    def clear_shard(self, shard):
        assert self.store.indices[shard].clear() is True
        for vid in [v for v, m in self.model.items() if m["shard"] == shard]:
            del self.model[vid]
            self.gone.add(vid)
            self.explicit_live.discard(vid)

    @rule()
    def clear_all(self):
        assert self.store.clear() == len(self.model)
        self.gone.update(self.model)
        self.model.clear()
        self.explicit_live.clear()
        self.used_prefixes.clear()

    @rule(other_devices=st.sampled_from([None, "0-1", "0-3"]))
    def save_and_reload(self, other_devices):
        assert self.store.save() is True
        self.store.close()
        keep, self.devices = self.devices, other_devices      # a store saved on G devices loads on G' (re-striped)
        try:
            self.store = self._open()
        finally:
            self.devices = keep
        # (the reloaded store keeps running on `other_devices` until the next reload: layouts are interchangeable)

    # ------------------------------------------------------------------ invariants
    def _expected(self, q, limit, flt=None, prefilter=False, threshold=0.0):
        """the reference's result list over the model's live rows"""
        live = sorted(self.model.items(), key=lambda kv: kv[1]["order"])
        if not live or limit <= 0:
            return []
        qn = q.astype(np.float64) / max(np.linalg.norm(q.astype(np.float64)), 1e-300)
        scored = []
        for vid, m in live:
            x = m["vec"].astype(np.float64)
            nx = np.linalg.norm(x)
            scored.append((vid, float(x @ qn / nx) if nx > 0 else 0.0, m))
        match = (lambda m: all(m["meta"].get(k) == v for k, v in flt.items())) if flt else (lambda m: True)
        if flt and not prefilter:
            # post-filter: one top-`limit` list per shard is the candidate set (vector_store.py:323-345)
            cand = []
            for s in range(S):
                in_s = [t for t in scored if t[2]["shard"] == s]
                k = min(limit, max(sum(1 for t in scored if t[2]["shard"] == ss) for ss in range(S)))
                cand += sorted(in_s, key=lambda t: -t[1])[:k]
            scored = cand
        if threshold > 0:     # after the merge, before the filter (vector_store.py:333-342); the pre-filter uses it as a floor
            scored = [t for t in scored if t[1] >= threshold]
        scored = [t for t in scored if match(t[2])]
        return [(vid, m["meta"]) for vid, _, m in sorted(scored, key=lambda t: -t[1])[:min(limit, len(live))]]

    @invariant()
    def agrees_with_model(self):
        if getattr(self, "store", None) is None:
            return
        st_ = self.store
        assert st_.count() == len(self.model)
        for vid in list(self.model)[:4] + sorted(self.gone)[:6] + sorted(self.gone)[-3:] + ["e_unknown", "p0_01"]:
            got = st_.get(vid)
            if vid in self.model:
                assert got is not None and np.array_equal(np.asarray(got[0], np.float32), self.model[vid]["vec"]) \
                    and got[1] == self.model[vid]["meta"], vid
            else:
                assert got is None, vid
        q = _vec(self.seed * 7 + 3)
        for limit in (1, 4, len(self.model) + 2):
            got = st_.search(q.tolist(), limit=limit)
            assert [(r[0], r[2]) for r in got] == self._expected(q, limit), ("search", limit)
            assert all(a[1] >= b[1] for a, b in zip(got, got[1:]))
        # threshold: applied after the merge, only when > 0 (vector_store.py:333-334); async and batch entry points
        full = st_.search(q.tolist(), limit=len(self.model) + 2)
        if full:
            t = sorted(r[1] for r in full)[len(full) // 2] - 1e-4
            want = [r for r in full if r[1] >= t] if t > 0 else full
            assert st_.search(q.tolist(), limit=len(self.model) + 2, threshold=t) == want, "threshold"
            assert asyncio.run(st_.search_async(q.tolist(), limit=4)) == full[:4], "async"
            br = st_.search_batch(np.stack([q, -q]), limit=3)
            assert br.ids()[0] == [r[0] for r in full[:3]] and br.ids()[1] == [r[0] for r in full[::-1][:3]], "batch"
        flt = {"g": 1}
        got = st_.search(q.tolist(), limit=3, filter_metadata=flt)
        assert [(r[0], r[2]) for r in got] == self._expected(q, 3, flt), "post-filter"
        st_.prefilter = True
        try:
            got = st_.search(q.tolist(), limit=3, filter_metadata=flt)
        finally:
            st_.prefilter = False
        assert [(r[0], r[2]) for r in got] == self._expected(q, 3, flt, prefilter=True), "pre-filter"
        if full:   # threshold together with a filter, both semantics (t sits between two scores: no boundary ties)
            ss = sorted((r[1] for r in full), reverse=True)
            hi, lo = ss[len(ss) // 3], ss[min(len(ss) // 3 + 1, len(ss) - 1)]
            t = (hi + lo) / 2 if hi - lo > 1e-4 else lo - 1e-3
            got = st_.search(q.tolist(), limit=3, threshold=t, filter_metadata=flt)
            assert [(r[0], r[2]) for r in got] == self._expected(q, 3, flt, threshold=t), "post-filter + threshold"
            st_.prefilter = True
            try:
                got = st_.search(q.tolist(), limit=3, threshold=t, filter_metadata=flt)
            finally:
                st_.prefilter = False
            assert [(r[0], r[2]) for r in got] == self._expected(q, 3, flt, prefilter=True, threshold=t), "pre-filter + threshold"


class MultiDeviceStoreMachine(StoreMachine):
    devices = "0-2"


import os  # noqa: E402

_settings = settings(max_examples=int(os.environ.get("WDBX_MODEL_EXAMPLES", "60")), stateful_step_count=int(os.environ.get("WDBX_MODEL_STEPS", "30")), deadline=None,
                     derandomize=os.environ.get("WDBX_MODEL_RANDOM", "") == "",   # the gate runs a fixed walk; explore with WDBX_MODEL_RANDOM=1
                     suppress_health_check=[HealthCheck.too_slow, HealthCheck.filter_too_much, HealthCheck.data_too_large])
TestStoreModel = StoreMachine.TestCase
TestStoreModel.settings = _settings
TestMultiDeviceStoreModel = MultiDeviceStoreMachine.TestCase
TestMultiDeviceStoreModel.settings = _settings
