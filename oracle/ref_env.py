"""TEST INFRASTRUCTURE ONLY -- makes the UNMODIFIED reference importable in the
build container: puts the timm shim (oracle/timm_shim) and /root/reference on
sys.path.  /root/reference does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this; it is used by
``oracle/make_golden.py`` and by the ``not gpu`` oracle-pinning tests only.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("SOCCDPT_REFERENCE_ROOT", "/root/reference")
SHIM_ROOT = os.path.join(HERE, "timm_shim")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "SOccDPT", "model"))


def enable_shim():
    if SHIM_ROOT not in sys.path:
        sys.path.insert(0, SHIM_ROOT)


def import_reference():
    """Returns the reference's ``SOccDPT.model`` sub-modules (loader, SOccDPT)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    enable_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import SOccDPT.model.loader as ref_loader  # noqa: E402
    import SOccDPT.model.SOccDPT as ref_model  # noqa: E402
    return ref_loader, ref_model
