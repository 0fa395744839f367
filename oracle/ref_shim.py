"""Import the UNMODIFIED reference -- from /root/reference in the build container, else from the git-ignored
verbatim install ``baseline/_ref`` (tools/install_reference.sh, run by ``__graft_entry__.build()``), which travels to
the GPU box with the gpurun snapshot.

TEST / BASELINE INFRASTRUCTURE.  Used by ``oracle/gen_golden.py`` (fixture generation), the ``reference``-marked tests
(they skip when no tree is available) and ``bench.py --impl reference``.  The product never imports it.

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

_INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")


def _resolve_root() -> str:
    env = os.environ.get("VSIQ_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _INSTALLED):
        if os.path.isdir(os.path.join(cand, "quantizers")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()

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
