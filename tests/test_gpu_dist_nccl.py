"""R-rank data-parallel correctness on GPUs: two NCCL ranks under torchrun against the
single-process update on the concatenated minibatch -- eager path and the three-segment CUDA-graph path with the
all-reduces between the graph replays.  The worker (tests/dist_worker_nccl.py) holds the assertions."""
import os
import re
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (run with `gpurun --gpus 2`)")
def test_two_rank_update_equals_the_single_process_update_on_the_concatenated_minibatch():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_worker_nccl.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    # the two ranks share one stdout: their report lines can land on one line, so count the reports, not the lines
    reports = ["DIST_REPORT " + body for body in re.findall(r"DIST_REPORT (\{.*?\})", res.stdout)]
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):  # kept next to the other run artefacts
        with open(os.path.join(out_dir, "dist_nccl_reports.txt"), "w") as f:
            f.write("\n".join(reports) + "\n")
    assert res.returncode == 0 and len(reports) == 2, ("\n".join(reports), res.stderr[-2000:])
