"""Loader of the unmodified reference from `baseline/_ref/` (see install_ref.py) for `bench.py`.

Only `bench.py --impl reference`, the `cpu_baseline` leg of `bench.py` and tests that compare
against the live reference use this; nothing under `dicp_b200/` does.  The reference's modules
import each other by the absolute name `dICP` (`ICP.py:9-10`), which in this repo is also the name
of the drop-in alias package, so the loader imports the reference with `baseline/_ref` first on
`sys.path` while any already-imported `dICP*` modules are set aside, and then restores them: both
implementations can live in one process.  `matplotlib` (imported at `ICP.py:8`, unused on the
compute path, not installed here) is stubbed with empty modules.
"""
import importlib
import os
import sys
import tempfile
import types

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_cached = None


def available():
    return os.path.exists(os.path.join(REF_DIR, "dICP", "ICP.py"))


def load_reference():
    """-> namespace with ICP, nn, loss of the unmodified reference, or None when baseline/_ref is absent."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        return None
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    aside = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "dICP" or k.startswith("dICP.")}
    sys.path.insert(0, REF_DIR)
    try:
        pkg = importlib.import_module("dICP")
        assert os.path.abspath(pkg.__file__).startswith(REF_DIR), pkg.__file__
        ns = types.SimpleNamespace(
            ICP=importlib.import_module("dICP.ICP").ICP,
            nn=importlib.import_module("dICP.nn").nn,
            loss=importlib.import_module("dICP.loss").loss,
            path=REF_DIR)
    finally:
        sys.path.remove(REF_DIR)
        for k in [k for k in sys.modules if k == "dICP" or k.startswith("dICP.")]:
            del sys.modules[k]
        sys.modules.update(aside)
    _cached = ns
    return ns


def reference_config(**overrides):
    """Path of a temporary copy of the reference's own YAML with some keys overridden (e.g. const_iter=True)."""
    with open(os.path.join(REF_DIR, "config", "dICP_config.yaml")) as f:
        cfg = yaml.safe_load(f)
    for key, value in overrides.items():
        hit = False
        for section in cfg["dICP"].values():
            if key in section:
                section[key] = value
                hit = True
        assert hit, "unknown config key %r" % key
    fd, path = tempfile.mkstemp(suffix=".yaml", prefix="dicp_ref_cfg_")
    with os.fdopen(fd, "w") as f:
        yaml.safe_dump(cfg, f)
    return path
