"""Import shim for the UNMODIFIED reference (ss0832/MultiOptPy) at /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (run in the build
container, where /root/reference exists) to produce the golden vectors under
``tests/golden/``.  Nothing in the product path, the ``-m gpu`` tests,
``smoke()`` or ``bench.py`` may import this module: /root/reference does not
exist on the GPU box.

``import multioptpy`` itself fails offline (matplotlib / ase / tblite are not
installed), so the package ``__init__`` is bypassed by registering an empty
namespace module whose ``__path__`` points at the reference tree (SURVEY.md
§8c / Appendix B.1).
"""
import importlib
import os
import sys
import types

REF_ROOT = os.environ.get("MOP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "multioptpy"))


def install():
    if "multioptpy" in sys.modules and getattr(sys.modules["multioptpy"], "_mop_shim", False):
        return sys.modules["multioptpy"]
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    pkg = types.ModuleType("multioptpy")
    pkg.__path__ = [os.path.join(REF_ROOT, "multioptpy")]
    pkg._mop_shim = True
    sys.modules["multioptpy"] = pkg
    return pkg


def ref(modname: str):
    """ref('Optimizer.rsirfo') -> the reference module multioptpy.Optimizer.rsirfo"""
    install()
    return importlib.import_module("multioptpy." + modname)
