"""Multi-GPU test of the fused NVLink reduce-scatter + clamp + RMSprop + all-gather kernel (csrc/dp_fused.cu): needs
>= 2 B200s on one box (skipped on a single-GPU box; the host-side reduction logic is covered on CPU by
tests/test_dp_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("size", ["small", "full"])
def test_fused_dp_step_two_or_more_ranks(size):
    import novel_vqa_b200 as nv
    n = nv.device_count()
    if n == 0:
        pytest.fail("no sm_100 device visible")
    if n < 2:
        pytest.skip("needs >= 2 GPUs on one box")
    world = min(n, 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "dp_fused_worker.py"), size]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DP_FUSED_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
