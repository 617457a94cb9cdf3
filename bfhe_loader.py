"""Registers the product package (whose directory name contains '-') as module ``bfhe_b200`` and the
test-only oracle wrapper as ``bfhe_oracle``.  Plumbing only."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "openfhe-boolean-circuit-evaluator_b200")


def _load(name, path, is_pkg=False):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, path, submodule_search_locations=[os.path.dirname(path)] if is_pkg else None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def load_package():
    return _load("bfhe_b200", os.path.join(PKG_DIR, "__init__.py"), is_pkg=True)


def load_build():
    return _load("bfhe_b200_build", os.path.join(PKG_DIR, "build.py"))


def load_oracle():
    """TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline)."""
    return _load("bfhe_oracle", os.path.join(ROOT, "oracle", "oracle.py"))
