"""Replay of tests/golden/api_probe_golden.json: the call sequences the generator ran against the reference's
unmodified VectorStore / WDBX (tests/golden/make_api_golden.py) must give the same outcomes -- return values,
exception types and messages -- through wdbx_b200, except at the steps the fixture lists as deliberate deviations.
CPU: numpy engine double; `-m gpu`: the real engine on a B200."""
import json
import tempfile
from pathlib import Path

import pytest

import wdbx_b200
from tests.fake_engine import FakeEngine

GOLDEN = Path(__file__).resolve().parent / "golden"


def _play():
    # the generator's `play` / `norm` helpers only (nothing in them touches /root/reference)
    src = (GOLDEN / "make_api_golden.py").read_text()
    ns = {}
    exec("import asyncio\n" + src[src.index("def norm(x):"):src.index("def main():")], ns)
    return ns["play"]


def _check(kind, script, want, got, deviations):
    assert len(got) == len(want) == len(script)
    for (name, method, args), w, g in zip(script, want, got):
        if name in deviations.get(kind, {}):
            ours, why = deviations[kind][name]
            assert g == ours, (kind, name, "deviation expected", ours, "got", g, why)
            assert g != w, (kind, name, "listed as a deviation but equals the reference")
        else:
            assert g == w, (kind, name, method, args, "reference", w, "ours", g)


def _run(factory_kw):
    g = json.loads((GOLDEN / "api_probe_golden.json").read_text())
    play = _play()
    with tempfile.TemporaryDirectory() as t1:
        st = wdbx_b200.VectorStore(4, t1, num_shards=2, **factory_kw)
        got = play(st, g["store_script"])
        st.close()
    _check("store", g["store_script"], g["store"], got, g["deviations"])
    return g, play


def test_vector_store_script_matches_the_reference():
    _run({"dist": wdbx_b200.DistContext(0, 1, 0), "_engine_factory": FakeEngine})


def _index_script(factory_kw):
    """the VectorIndex operator boundary: the reference's FaissIndex vs the shard facade of a one-shard store"""
    g = json.loads((GOLDEN / "api_probe_golden.json").read_text())
    with tempfile.TemporaryDirectory() as t0:
        st = wdbx_b200.VectorStore(4, t0, num_shards=1, **factory_kw)
        got = _play()(st.indices[0], g["index_script"])
        st.close()
    _check("index", g["index_script"], g["index"], got, g["deviations"])


def test_vector_index_script_matches_the_reference():
    _index_script({"dist": wdbx_b200.DistContext(0, 1, 0), "_engine_factory": FakeEngine})


def test_facade_script_matches_the_reference(monkeypatch):
    import wdbx_b200.vector_store as vsmod

    g = json.loads((GOLDEN / "api_probe_golden.json").read_text())
    orig = vsmod.VectorStore.__init__

    def patched(self, *a, **kw):   # the facade builds its own store: give that one the numpy double too
        kw.setdefault("_engine_factory", FakeEngine)
        kw.setdefault("dist", wdbx_b200.DistContext(0, 1, 0))
        orig(self, *a, **kw)

    monkeypatch.setattr(vsmod.VectorStore, "__init__", patched)
    with tempfile.TemporaryDirectory() as t2:
        got = _play()(wdbx_b200.WDBX(vector_dimension=4, num_shards=2, data_dir=t2, log_level="ERROR"), g["facade_script"])
    _check("facade", g["facade_script"], g["facade"], got, g["deviations"])


@pytest.mark.gpu
def test_scripts_match_the_reference_on_the_device(built_lib):
    _index_script({})
    g, play = _run({})
    with tempfile.TemporaryDirectory() as t2:
        got = play(wdbx_b200.WDBX(vector_dimension=4, num_shards=2, data_dir=t2, log_level="ERROR"), g["facade_script"])
    _check("facade", g["facade_script"], g["facade"], got, g["deviations"])


def test_public_surface_matches_the_reference():
    """Every public method of the reference's VectorIndex ABC / FaissIndex / VectorStore / WDBX facade exists here with
    the same parameter names, the same simple defaults and the same sync / async nature (the fixture lists them as
    inspected on the reference's classes).  Not carried over, by scope: the plugin registry of the
    facade; `WDBX.vector_store` is the store attribute, which shadows the method of that name in the reference too."""
    import inspect

    from wdbx_b200.indexing import B200FlatIndex, VectorIndex

    g = json.loads((GOLDEN / "api_probe_golden.json").read_text())["surface"]
    ours = {"VectorIndex": VectorIndex, "FaissIndex": B200FlatIndex, "VectorStore": wdbx_b200.VectorStore, "WDBX": wdbx_b200.WDBX}
    out_of_scope = {"WDBX": {"get_plugin", "register_plugin", "vector_store"}}
    assert sorted(VectorIndex.__abstractmethods__) == g["VectorIndex.abstract"]
    checked = 0
    for cls_name, methods in g.items():
        if cls_name not in ours:
            continue
        for name, want in methods.items():
            if name in out_of_scope.get(cls_name, ()):
                continue
            f = getattr(ours[cls_name], name, None)
            assert callable(f), f"{cls_name}.{name} is missing"
            sig = inspect.signature(f)
            params = [p for p in sig.parameters if p != "self"]
            if want["params"]:                   # (the reference's VectorStore.optimize_async inspects as "()": nothing to compare)
                assert params == want["params"], (cls_name, name, params, want["params"])
            for k, v in want["defaults"].items():
                if not (cls_name == "WDBX" and name == "__init__"):
                    assert sig.parameters[k].default == v, (cls_name, name, k)
            assert inspect.iscoroutinefunction(f) == want["async"], (cls_name, name)
            checked += 1
    assert checked >= 60
