"""Import the UNMODIFIED reference from /root/reference -- build-container only.

TEST INFRASTRUCTURE.  /root/reference does not exist on the GPU box, so nothing
that runs there (``-m gpu`` tests, smoke(), bench.py) may import this module;
it is used by ``oracle/gen_golden.py`` (fixture generation) and by the
``reference``-marked CPU tests that skip when the tree is absent.

The reference's hot-path modules carry three junk imports that no longer
resolve on Python 3.12 (``import imp`` -- quantizers/fake_quantize.py:1,
modules/fused.py:1; ``from tkinter import W`` -- modules/fused.py:2;
``from turtle import forward`` -- modules/fused.py:3).  They are unused, so
three dummy ``sys.modules`` entries make the tree importable without touching
any reference file.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VSIQ_REFERENCE_ROOT", "/root/reference")

# Top-level package names the reference uses (implicit namespace packages).
_REF_PACKAGES = ("utils", "observers", "quantizers", "modules", "nets", "dataset")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "quantizers"))


def install() -> None:
    """Put the reference on sys.path (front) with the junk-import shim."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    for name, attrs in (("imp", {}), ("tkinter", {"W": "w"}),
                        ("turtle", {"forward": lambda *a, **k: None})):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def uninstall() -> None:
    """Drop the reference's modules again so they cannot shadow anything else."""
    if REFERENCE_ROOT in sys.path:
        sys.path.remove(REFERENCE_ROOT)
    for k in list(sys.modules):
        if k.split(".")[0] in _REF_PACKAGES:
            f = getattr(sys.modules[k], "__file__", None) or ""
            p = getattr(sys.modules[k], "__path__", None)
            if f.startswith(REFERENCE_ROOT) or (p is not None and any(str(q).startswith(REFERENCE_ROOT) for q in p)):
                del sys.modules[k]
