"""Slab-partitioned (one process per GPU, NCCL) solve vs the oracle -- needs >= 2 GPUs, skipped otherwise (never two ranks on one GPU)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "multi_gpu_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "MULTI_GPU_PARITY_OK" in out.stdout


def _torchrun(n, script, args, env=None, port="29521", timeout=900):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
           "--master-port", port, os.path.join(ROOT, "tests", script)] + args
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=e)


@pytest.mark.parametrize("nccl_only", [False, True])
def test_ranks_agree_with_one_gpu(tmp_path, nccl_only):
    """2 (and 4, when present) ranks vs the same problems on ONE GPU: interior + ghost-class tiles, fused iteration with the halo exchange on the
    second stream, ghost planes of 17.7 k doubles (multi-block peer-memory halo kernel); with the peer-memory exchange and with NCCL only."""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    ref = str(tmp_path / "ref.npz")
    cases = "diph3d" if nccl_only else "diph3d,mono3d,diph2d,diph2d_heads"
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "multi_gpu_variants.py"), "--single", ref, "--cases", cases], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "SINGLE_GPU_REFERENCE_WRITTEN" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    for n in ([2, 4] if ng >= 4 and not nccl_only else [2]):
        env = {"PB200_REPORT_HEADS": "1"}
        if nccl_only:
            env["PB200_NO_P2P"] = "1"
        out = _torchrun(n, "multi_gpu_variants.py", ["--ref", ref, "--cases", cases], env=env)
        assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
        assert "MULTI_GPU_VARIANTS_OK" in out.stdout
        # (diph2d_heads is the geometry for which the band heads could run with several ranks -- PB200_BANDFUSE_MULTI=1; off by default, see fold_build)


def test_one_process_drives_two_gpus(tmp_path):
    """pb200_init_multi: an unchanged single-process script (global host arrays) on 2 GPUs vs the same script on one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ref = str(tmp_path / "ref.npz")
    script = os.path.join(ROOT, "tests", "multi_gpu_variants.py")
    out = subprocess.run([sys.executable, script, "--single", ref, "--cases", "diph3d,mono3d"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    out = subprocess.run([sys.executable, script, "--team", "2", "--ref", ref, "--cases", "diph3d,mono3d"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "SINGLE_PROCESS_MULTI_GPU_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
