"""B200-native GraphSAGE minibatch hot path behind the class API of the reference's
`src/models.py`.  See DESIGN.md.  Importing this package never imports `oracle/`."""
import importlib as _importlib

__version__ = "0.1.0"

from . import synth  # noqa: F401  (numpy only)

_LAZY_CLASSES = ("GraphSage", "SageLayer", "Classification", "UnsupervisedLoss")
_LAZY_MODULES = ("models", "native", "ops", "graph", "trainer", "build", "inference", "peer", "datacache")


def __getattr__(name):
    # torch-dependent modules are loaded lazily so `import graphsage_b200.synth` stays light
    if name in _LAZY_CLASSES:
        return getattr(_importlib.import_module(__name__ + ".models"), name)
    if name in _LAZY_MODULES:
        return _importlib.import_module(__name__ + "." + name)
    if name == "AdjCSR":
        return _importlib.import_module(__name__ + ".graph").AdjCSR
    raise AttributeError(name)
