"""Importable alias of the package directory `qp-plonky2_b200/` (a hyphen cannot appear in a
Python module name).  Everything lives there; this file only re-exports it."""
import importlib.util
import os
import sys

_real = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qp-plonky2_b200")
_spec = importlib.util.spec_from_file_location(
    "_qp_plonky2_b200_impl", os.path.join(_real, "__init__.py"), submodule_search_locations=[_real]
)
_impl = importlib.util.module_from_spec(_spec)
sys.modules["_qp_plonky2_b200_impl"] = _impl
_spec.loader.exec_module(_impl)
globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
__path__ = [_real]  # submodules (qp_plonky2_b200.dist, ...) resolve inside qp-plonky2_b200/
