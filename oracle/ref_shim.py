"""Import shim for running the UNMODIFIED reference (/root/reference) on CPU.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` in the build container to
produce the committed fixtures under ``tests/golden/``; it is never imported by the
product package and /root/reference does not exist on the GPU box.

Two shims are needed:
  1. ``fairseq`` is not installed and reference ``model.py:10`` imports
     ``fairseq.data.Dictionary`` -> a stub package with an empty ``Dictionary``.
  2. the reference's ``datasets/`` directory (no ``__init__.py``) is shadowed by the
     installed HuggingFace ``datasets`` -> pre-seed ``sys.modules['datasets']`` with a
     namespace module whose ``__path__`` is the reference directory.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MH_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def install():
    """Make ``import model, module, upstream...`` resolve to the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if "fairseq" not in sys.modules:
        fairseq = types.ModuleType("fairseq")
        fairseq.__path__ = []
        data = types.ModuleType("fairseq.data")

        class Dictionary:  # reference model.py:10 only needs the name
            pass

        data.Dictionary = Dictionary
        fairseq.data = data
        sys.modules["fairseq"] = fairseq
        sys.modules["fairseq.data"] = data
    ds = types.ModuleType("datasets")
    ds.__path__ = [os.path.join(REFERENCE_ROOT, "datasets")]
    sys.modules["datasets"] = ds
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_yaml(rel_path):
    import yaml

    with open(os.path.join(REFERENCE_ROOT, rel_path)) as f:
        return yaml.load(f, Loader=yaml.FullLoader)
