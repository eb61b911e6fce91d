"""TEST INFRASTRUCTURE ONLY -- imports the unmodified reference (build container only).

/root/reference does not exist on the GPU box; every caller must handle
`reference_available() == False` (tests skip, goldens are read from tests/golden).
Recipe = SURVEY.md Appendix B.
"""
import os
import sys

REF_ROOT = "/root/reference"
SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")


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
    import sam.sa_m4c as ref_model
    import sam.spatial_utils as ref_spatial
    return ref_model, ref_spatial, registry
