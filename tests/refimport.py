"""Import the UNMODIFIED reference from baseline/_ref (scripts/install_ref.py put it there; it travels to the
GPU box with the gpurun snapshot) with vael_b200.compat standing in for `problog`.  The reference's task /
plot modules import matplotlib, seaborn and scikit-image at module level; those are absent from the image and
irrelevant to the logic path, so inert stand-ins are registered for whichever of them is missing."""
import hashlib
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
MANIFEST = os.path.join(ROOT, "tests", "golden", "reference_manifest.json")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "models", "vael.py"))


def modified_files():
    """Files under baseline/_ref whose sha256 differs from the committed manifest of the reference."""
    want = json.load(open(MANIFEST))["files"]
    bad = []
    for rel, sha in want.items():
        p = os.path.join(REF_DIR, rel)
        if not os.path.exists(p) or hashlib.sha256(open(p, "rb").read()).hexdigest() != sha:
            bad.append(rel)
    return bad


class _Inert(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Inert(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return None


def install():
    """compat.install() + sys.path; returns nothing.  Safe to call twice."""
    from vael_b200 import compat
    compat.install(force=True)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "seaborn", "skimage", "skimage.transform"):
        if name in sys.modules:
            continue
        try:
            __import__(name)
        except ImportError:
            sys.modules[name] = _Inert(name)
    # the oracle's golden generator may have imported same-named modules (models, utils) from elsewhere
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    for name in list(sys.modules):
        if name.split(".")[0] in ("models", "utils", "config", "const_define"):
            f = getattr(sys.modules[name], "__file__", None) or ""
            if f and not f.startswith(REF_DIR):
                del sys.modules[name]
