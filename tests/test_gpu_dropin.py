"""The boundary the reference's trainer actually uses: `sam.sa_m4c` swapped for the samk module, the UNMODIFIED
`sam/task_utils.py` (forward_model, clip_gradients, get_optim_scheduler) driving it.  Runs tests/dropin_train_smoke.py
in a fresh interpreter (the alias has to be in place before anything imports `sam.*`)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("precision,tol", [("f16", 2e-3), ("bf16x3", 2e-4)])
def test_reference_train_loop_drives_the_samk_module(precision, tol):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_train_smoke.py"), "--iters", "3",
                        "--precision", precision], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    out = json.loads(r.stdout.strip().splitlines()[-1])
    if "skipped" in out:
        pytest.skip(out["skipped"])
    assert out["precision"] == precision and out["kernels_launched"] > 100
    # three optimizer updates through the reference's own loop: the loss curves of the two models stay together
    for a, b in zip(out["loss_samk"], out["loss_reference"]):
        assert abs(a - b) <= tol * abs(b), out
    assert out["loss_samk"][-1] < out["loss_samk"][0]              # and the updates did train
    assert out["acc_samk"] == out["acc_reference"] and out["predictions_equal"]
