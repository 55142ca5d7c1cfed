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


def _install_repaired_vit_module():
    """The reference's hybrid constructor raises ``NameError: name 'value' is not defined``
    (SOccDPT/model/backbones/vit.py:181-182, 222-223: the Sequential is bound to ``_`` but ``exec`` assigns
    ``value``; the intended line survives as a comment right above).  The one-token repair is applied to the
    module's source text IN MEMORY before it is imported -- nothing of the reference is copied into this repo."""
    import importlib.util
    import types
    name = "SOccDPT.model.backbones.vit"
    if name in sys.modules:
        return
    import SOccDPT.model.backbones  # noqa: F401  (package object; its __init__ is empty)
    path = os.path.join(REFERENCE_ROOT, "SOccDPT", "model", "backbones", "vit.py")
    with open(path) as f:
        src = f.read()
    patched = src.replace("        _ = nn.Sequential(", "        value = nn.Sequential(")
    assert patched != src and patched.count("value = nn.Sequential(") >= 2, "reference vit.py no longer matches the repair"
    mod = types.ModuleType(name)
    mod.__file__ = path
    mod.__package__ = "SOccDPT.model.backbones"
    sys.modules[name] = mod
    exec(compile(patched, path, "exec"), mod.__dict__)
    setattr(sys.modules["SOccDPT.model.backbones"], "vit", mod)


def import_reference(repair_hybrid=True):
    """Returns the reference's ``SOccDPT.model`` sub-modules (loader, SOccDPT).
    The hybrid repair (see above) is installed by default; it only touches the two broken lines."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    enable_shim()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if repair_hybrid:
        _install_repaired_vit_module()
    import SOccDPT.model.loader as ref_loader  # noqa: E402
    import SOccDPT.model.SOccDPT as ref_model  # noqa: E402
    return ref_loader, ref_model


def load_reference_function(rel_path, name, namespace=None):
    """Compiles ONE top-level function of a reference source file in memory (for modules whose own imports -- matplotlib,
    wandb -- are missing here).  The text is read from /root/reference at run time and never written into this repo."""
    import ast
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path) as f:
        src = f.read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            ns = dict(namespace or {})
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns[name]
    raise KeyError(f"{name} not found in {path}")


def load_reference_class(rel_path, name, namespace=None):
    """Compiles ONE top-level class of a reference source file in memory (same purpose and rules as load_reference_function)."""
    import ast
    path = os.path.join(REFERENCE_ROOT, rel_path)
    with open(path) as f:
        src = f.read()
    for node in ast.parse(src).body:
        if isinstance(node, ast.ClassDef) and node.name == name:
            ns = dict(namespace or {})
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns[name]
    raise KeyError(f"{name} not found in {path}")
