"""Import the vendored REAL reference (oracle/_ref, built by oracle/build_ref.py)  --  TEST /
BENCH INFRASTRUCTURE, never imported by the product (`uglad_b200`).

    ref = load()          # None when oracle/_ref has not been built
    ref.main.forward_uGLAD(...), ref.glad.glad(...), ref.GladParams, ref.prepare_data, ref.metrics

matplotlib and pyvis (plotting only; absent from this image) get import-only stand-ins whose
every attribute is a no-op callable, so `uglad.main` imports and `plot_loss_curve` does nothing.
Nothing on the numeric path is touched.
"""
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


class _Noop(types.ModuleType):
    """Module whose attributes are all no-op callables (classes included: calling returns self)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _noop


def _noop(*a, **k):
    return _NOOP_OBJ


class _NoopObj:
    def __getattr__(self, name):
        return _noop

    def __call__(self, *a, **k):
        return self

    def __iter__(self):
        return iter(())


_NOOP_OBJ = _NoopObj()


def _stand_in(name, children=()):
    try:
        __import__(name)
        return
    except Exception:
        pass
    mod = _Noop(name)
    mod.__path__ = []
    sys.modules[name] = mod
    for c in children:
        sub = _Noop(f"{name}.{c}")
        sys.modules[f"{name}.{c}"] = sub
        setattr(mod, c, sub)


class Reference:
    def __init__(self, main, glad, GladParams, prepare_data, metrics):
        self.main, self.glad, self.GladParams, self.prepare_data, self.metrics = main, glad, GladParams, prepare_data, metrics


_cached = None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "uglad", "main.py"))


def load():
    """The vendored reference's modules, or None when oracle/_ref is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    _stand_in("matplotlib", ("colors", "pyplot"))
    _stand_in("pyvis", ("network",))
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import uglad.main as ref_main
    from uglad.glad import glad as ref_glad
    from uglad.glad.glad_params import GladParams
    from uglad.utils import metrics, prepare_data
    if not os.path.abspath(ref_main.__file__).startswith(REF_DIR):
        raise RuntimeError(f"'uglad' resolved to {ref_main.__file__}, not the vendored copy under {REF_DIR}")
    _cached = Reference(ref_main, ref_glad, GladParams, prepare_data, metrics)
    return _cached
