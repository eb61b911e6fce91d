"""TEST INFRASTRUCTURE ONLY -- imports the unmodified reference.

In the build container that is /root/reference; on the GPU box (where that path does not exist) it is the verbatim,
git-ignored copy oracle/_ref/ made by `python -m oracle.build_ref` (run by `__graft_entry__.build()`).  Every caller
must handle `reference_available() == False` (tests skip, goldens are read from tests/golden).
Recipe = SURVEY.md Appendix B.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(HERE, "shim")


def _ref_root():
    for root in ("/root/reference", os.path.join(HERE, "_ref")):
        if os.path.isfile(os.path.join(root, "sam", "sa_m4c.py")):
            return root
    return "/root/reference"


REF_ROOT = _ref_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "sam", "sa_m4c.py"))


def load_reference(vocab_size=5000):
    """Returns (sa_m4c module, spatial_utils module, registry) of the unmodified reference."""
    if not reference_available():
        raise RuntimeError("reference tree %s not present on this machine" % REF_ROOT)
    for p in (REF_ROOT, SHIM_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)
    from tools.registry import registry  # /root/reference/tools/registry.py:1-3
    registry.answer_vocab = ["w%d" % i for i in range(vocab_size)]
    registry.BOS_IDX, registry.EOS_IDX, registry.PAD_IDX = 1, 2, 0
    if "sam.phoc" not in sys.modules:      # the PHOC C extension (dataset side) is not built here; the hot path never calls it
        import types
        stub = types.ModuleType("sam.phoc")
        stub.build_phoc = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("PHOC extension not built (oracle shim)"))
        sys.modules["sam.phoc"] = stub
    import sam.sa_m4c as ref_model
    import sam.spatial_utils as ref_spatial
    return ref_model, ref_spatial, registry
