"""Stage the UNMODIFIED reference sources under baseline/_ref/ so they travel to the GPU box.

    python baseline/make_ref.py            # in the build container, where /root/reference exists

`baseline/_ref/` is git-ignored (the reference's text never enters this repository's history) but NOT
gpurun-ignored, so the copy ships with every gpurun snapshot.  Two users, both test / measurement
infrastructure -- nothing under graphsage-pytorch_b200/ imports it:

* `bench.py --impl reference` runs the reference's own `src/models.py` classes on the host cores
  (`cpu_baseline.kind = "reference"`), with the `random.sample` shim of oracle/ref_harness.py
  (Python >= 3.11 refuses a set population; CPython <= 3.10, which the reference targets, converted it
  with tuple() -- same draws);
* `tests/test_gpu_reference_loop.py` imports the reference's own `src/utils.py` and runs its `apply_model`
  and `evaluate` UNCHANGED against the drop-in classes of this package (the "drop-in test").

`pip install` of the reference (the base contract's recipe) does not apply: it has no setup.py /
pyproject.toml, it is a directory of scripts.  A digest of every staged file is written beside the copy so a
test can tell a stale or edited copy from the real thing.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("src/models.py", "src/utils.py", "src/dataCenter.py", "src/main.py", "src/experiments.conf", "README.md")


def stage(reference_root: str = "/root/reference") -> str:
    if not os.path.isfile(os.path.join(reference_root, "src", "models.py")):
        raise RuntimeError(f"no reference at {reference_root}")
    digests = {}
    for rel in FILES:
        src = os.path.join(reference_root, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as fp:
            digests[rel] = hashlib.sha256(fp.read()).hexdigest()
    with open(os.path.join(DEST, "src", "__init__.py"), "a"):     # `import src.models`, as the reference's main.py does
        pass
    with open(os.path.join(DEST, "DIGESTS.json"), "w") as fp:
        json.dump({"source": reference_root, "sha256": digests}, fp, indent=1)
    return DEST


def available() -> bool:
    return os.path.isfile(os.path.join(DEST, "src", "models.py"))


def verify() -> bool:
    """The staged files still have the digests recorded when they were copied."""
    try:
        want = json.load(open(os.path.join(DEST, "DIGESTS.json")))["sha256"]
    except Exception:
        return False
    for rel, d in want.items():
        try:
            with open(os.path.join(DEST, rel), "rb") as fp:
                if hashlib.sha256(fp.read()).hexdigest() != d:
                    return False
        except OSError:
            return False
    return bool(want)


def load(module: str = "models"):
    """Import `src.<module>` of the staged reference (with the random.sample shim installed first)."""
    if not available():
        raise RuntimeError("baseline/_ref is not staged: run `python baseline/make_ref.py` in the build container")
    import importlib
    import random
    if not getattr(random.sample, "_gsage_shim", False):
        original = random.sample

        def sample(population, k, **kw):
            if isinstance(population, (set, frozenset)):
                population = tuple(population)          # what CPython <= 3.10 did inside random.sample
            return original(population, k, **kw)

        sample._gsage_shim = True
        random.sample = sample
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    cached = sys.modules.get("src")
    if cached is not None and not os.path.abspath(getattr(cached, "__file__", "") or "").startswith(DEST):
        for name in [m for m in sys.modules if m == "src" or m.startswith("src.")]:
            del sys.modules[name]
    return importlib.import_module(f"src.{module}")


if __name__ == "__main__":
    print(stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference"))
