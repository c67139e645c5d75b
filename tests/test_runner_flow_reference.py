"""CPU, build container only (needs the reference checkout): the UNMODIFIED runners/evaluation_single.py — inference_pose and
inference_energy with their pickle hand-off — executed end to end over a synthetic detection pickle with genpose_b200 dropped in
by PYTHONPATH and the C library replaced by a numpy stand-in at the C ABI (tests/harness/runner_flow.py).  BASELINE configs[4]
(REAL275 accuracy) cannot be measured offline; this is the part of it that can: "drops in unchanged" is executed, not inferred."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GENPOSE_REFERENCE_ROOT", "/root/reference")


@pytest.mark.reference
@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "runners", "evaluation_single.py")), reason="reference tree not present")
def test_unmodified_runner_control_flow_on_our_agent(tmp_path):
    env = dict(os.environ, PYTHONPATH="", OMP_NUM_THREADS="4")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "harness", "runner_flow.py"), REF, ROOT, str(tmp_path)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert res.returncode == 0 and "RUNNER_FLOW_OK 9 6 3 3" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
