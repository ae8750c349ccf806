"""bench.py must print its headline line even when one of the extra configurations (other_configs) fails: each
is guarded and its failure becomes that configuration's record.  No GPU needed: the contexts are made to fail."""
import importlib.util
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("_bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_default_extras_depend_on_world_size(monkeypatch):
    bench = _load_bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    assert bench.parse().extra == "c2f,c3,c4,c5"
    monkeypatch.setenv("WORLD_SIZE", "8")
    assert bench.parse().extra == "c5"           # the configuration BASELINE.json names for 8 GPUs
    monkeypatch.setattr(sys, "argv", ["bench.py", "--extra", ""])
    assert bench.parse().extra == ""


def test_failing_extra_config_becomes_a_record(monkeypatch, capsys):
    bench = _load_bench()
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    a = bench.parse()
    closed = []

    class Boom(RuntimeError):
        pass

    class FakeCtx:
        def __init__(self, *args, **kw):
            pass

        def close(self):
            closed.append(1)

        def add_space(self, *args, **kw):
            raise Boom("no device")

    def fail(*args, **kw):
        raise Boom("no device")

    fake = types.SimpleNamespace(
        _lib=types.SimpleNamespace(load=lambda: types.SimpleNamespace()),
        B200Context=FakeCtx,
        B200CSR=types.SimpleNamespace(stencil_free=fail, stencil=fail),
        cgs2=None, mgs2=None, cgs=None, mgs=None, mgs2b=None)
    out = bench.other_configs(fake, a, 0, 1, 0, None)
    assert set(out) == {"c2_matrix_free", "c3", "c4", "c5"}
    for rec in out.values():
        assert rec["ok"] is False and "Boom: no device" in rec["error"]
    assert len(closed) == 4                     # every failed configuration gave its context back
    assert "Boom" in capsys.readouterr().err    # the traceback goes to stderr, the JSON line stays clean
