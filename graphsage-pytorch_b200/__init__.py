"""B200-native GraphSAGE minibatch hot path behind the class API of the reference's
`src/models.py`.  See DESIGN.md.  Importing this package never imports `oracle/`."""
__version__ = "0.1.0"

from . import synth  # noqa: F401  (numpy only)


def __getattr__(name):
    # torch-dependent modules are loaded lazily so `import graphsage_b200.synth` stays light
    if name in ("GraphSage", "SageLayer", "Classification", "UnsupervisedLoss", "models"):
        from . import models as _m
        return _m if name == "models" else getattr(_m, name)
    if name in ("native", "AdjCSR", "trainer"):
        import importlib
        mod = {"native": "native", "AdjCSR": "graph", "trainer": "trainer"}[name]
        m = importlib.import_module(f"{__name__}.{mod}")
        return getattr(m, name) if name == "AdjCSR" else m
    raise AttributeError(name)
