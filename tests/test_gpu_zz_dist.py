"""Row-sharded parity through tools/dist_check.py under torchrun: with one rank per GPU over NCCL + the NVLink
peer window when the box has >= 2 GPUs, and ALWAYS with two ranks sharing GPU 0 (peer window only, gloo for the
test's own gathers) so that a single-GPU box exercises the sharded path too.

(The file name sorts last on purpose: these are the slow multi-process tests — minutes of CPU oracle work per case —
and a `pytest -x` run should have been through every single-process GPU test before it gets here.  The round-2 GPU
scripts under tools/ still name the file as it was called then, tests/test_gpu_dist.py.)"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for ln in out.splitlines() if ln.startswith("GPU "))
    except (OSError, subprocess.TimeoutExpired):
        return 0


def _run(port, extra_env, timeout=600, nproc=2):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dist_check.py")]
    env = dict(os.environ, **extra_env)
    # own process group: on a timeout the launcher AND its workers (whose kernels may be spinning on a flag that
    # never comes) are killed together, so a bug cannot leave the GPU busy behind the test
    import signal
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=ROOT, env=env,
                            start_new_session=True)
    try:
        out, err = proc.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, err = proc.communicate()
        raise AssertionError("dist_check timed out\n" + out[-3000:] + err[-3000:])
    assert proc.returncode == 0, out[-3000:] + err[-3000:]
    assert "dist_check ok" in out


@pytest.mark.skipif(_ngpus() < 2, reason="needs at least 2 GPUs")
def test_row_sharded_eigsolve_two_ranks():
    _run(29611, {})


def test_row_sharded_two_ranks_on_one_gpu():
    """Two processes on GPU 0: CUDA-IPC peer windows, no NCCL.  Every cross-rank wait is resolved by the
    driver's time slicing between the two contexts, so this is slow per step but exercises the same kernels."""
    _run(29613, {"B2K_ONE_GPU": "1", "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]})


def test_row_sharded_three_ranks_on_one_gpu():
    """Three ranks on GPU 0: the middle rank has TWO neighbours (halo rows pushed both ways by the Gram-Schmidt
    launch, three-way rank-ordered sums) and the shards are unequal (151 grid lines over 3 ranks).  Primitives, the
    converged eigsolve and the device-chained fixed-cycle job against the serial oracle (the short form of dist_check)."""
    _run(29615, {"B2K_ONE_GPU": "1", "DIST_CHECK_SHORT": "1",
                 "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]}, nproc=3)


def test_peer_wait_watchdog():
    """A rank that drops out must not leave the others spinning inside a kernel: the in-kernel waits of the peer
    window give up after B2K_PEER_TIMEOUT_S and the call fails with B2K_ENCCL (tools/dist_check.py: watchdog_check)."""
    _run(29617, {"B2K_ONE_GPU": "1", "DIST_CHECK_WATCHDOG": "1",
                 "CUDA_VISIBLE_DEVICES": os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0]}, timeout=240)
