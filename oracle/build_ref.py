"""TEST INFRASTRUCTURE ONLY -- puts a verbatim copy of the reference's Python tree under oracle/_ref/.

    python -m oracle.build_ref

The reference is pure Python (nothing to compile), and /root/reference does not exist on the GPU box: the copy is
what lets the UNMODIFIED reference run there -- as the CPU arm of bench.py (`--impl reference`, `cpu_baseline`) and in
the drop-in test that drives the reference's own `forward_model` / `get_optim_scheduler` (sam/task_utils.py) against
the samk module.  oracle/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored.
Only `sam/` and `tools/` (*.py) are copied: train.py / evaluator.py need datasets that do not exist here.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")


def build():
    """Copy (or refresh) the tree; returns the destination, or None when the reference is not on this machine."""
    if not os.path.isfile(os.path.join(SRC, "sam", "sa_m4c.py")):
        return DST if os.path.isdir(DST) else None
    for top in ("sam", "tools"):
        for root, _dirs, files in os.walk(os.path.join(SRC, top)):
            rel = os.path.relpath(root, SRC)
            for f in files:
                if not f.endswith(".py"):
                    continue
                os.makedirs(os.path.join(DST, rel), exist_ok=True)
                s, d = os.path.join(root, f), os.path.join(DST, rel, f)
                if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
                    shutil.copyfile(s, d)
    return DST


if __name__ == "__main__":
    print(build())
