"""TEST INFRASTRUCTURE ONLY — import the reference's hot-path files UNMODIFIED.

Loads /root/reference/src/spVIPES/{module/spVIPESmodule.py, nn/networks.py, nn/utils.py}
in place (nothing is copied) against the scvi stand-in in oracle/scvi_stub.  The reference's
package __init__ (src/spVIPES/__init__.py:9) imports anndata/scanpy/rich-dependent
subpackages that are absent offline, so a bare namespace package is registered instead and
only the hot-path submodules are executed.

Only usable in the authoring container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py and by the `-m "not gpu"` tests that validate oracle/restatement.py.
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("SPVIPES_REFERENCE", "/root/reference")
_SRC = os.path.join(REF_ROOT, "src", "spVIPES")
_STUB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "scvi_stub")


def available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "module", "spVIPESmodule.py"))


def load():
    """Return (spVIPESmodule class, networks module) from the unmodified reference."""
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    if _STUB not in sys.path:
        sys.path.insert(0, _STUB)
    if "spVIPES.module.spVIPESmodule" not in sys.modules:
        for name, sub in (("spVIPES", ""), ("spVIPES.nn", "nn"), ("spVIPES.module", "module")):
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(_SRC, sub) if sub else _SRC]
            m.__package__ = name
            sys.modules[name] = m
        importlib.import_module("spVIPES.nn.utils")
        importlib.import_module("spVIPES.nn.networks")
        importlib.import_module("spVIPES.module.spVIPESmodule")
    return sys.modules["spVIPES.module.spVIPESmodule"].spVIPESmodule, sys.modules["spVIPES.nn.networks"]
