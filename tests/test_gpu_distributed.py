"""Multi-GPU parity under -m gpu: spawns 2 ranks (torchrun) when >= 2 GPUs are visible, skips otherwise. Each case runs
the distributed solve of tools/dist_solve.py -- C++ host layer, NCCL communicator, owner-push halo exchange -- with
--check, which makes rank 0 also solve the same global problem on one GPU and assert identical iteration / cycle counts
(the reference's distributed-equals-serial criterion, frontend/tests/mpi/sfem_ParallelLaplacianApplyTest.cpp:133-141,
extended to CG and multigrid)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, args, timeout=600):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, {torch.cuda.device_count()} visible")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_solve.py")] + args
    env = dict(os.environ, B200FEM_COMM_TIMEOUT_S="60")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(lines[-1])


def test_distributed_cg_tet4_matches_serial():
    out = _run(2, ["c4", "--side", "33", "--rtol", "1e-8", "--check"])
    assert out["rc"] == 0 and out["check"]["serial_iterations"] == out["iterations"]
    assert out["check"]["max_rel_residual_history_diff"] < 1e-6, out["check"]
    assert out["halo"]["transport"] in ("push", "nccl")


def test_distributed_mg_elasticity_matches_serial():
    out = _run(2, ["c5", "--macro", "6", "--layers", "3", "--level", "4", "--sweeps", "10", "--check"])
    assert out["rc"] == 0 and out["check"]["serial_cycles"] == out["cycles"]
    assert out["residual_history"][-1] < 1e-9


def test_overlapped_apply_matches_plain_distributed_apply():
    """the ghost-independent elements between the two halves of the halo gather (one stream, sub-range launches of the
    warp-specialised kernel) give the rows the gather-then-apply form gives"""
    out = _run(2, ["apply", "--macro", "8", "--layers", "4", "--level", "8"])
    assert out["rc"] == 0 and out["overlapped"] is True, out
    assert out["max_rel_diff"] < 1e-13, out
    # the exchange restricted to the ghosts that can reach an owned row, the other ghost entries poisoned
    assert out["max_rel_diff_reduced_halo"] < 1e-13, out
    assert out["values_sent_rank0_reduced"] <= out["values_sent_rank0"]


def test_single_rank_goes_through_the_same_layer():
    """world 1: Communicator / Halo / ParallelMatrixFreeOperator degenerate to the plain solve (no NCCL call)."""
    out = _run(1, ["c4", "--side", "25", "--check"])
    assert out["rc"] == 0 and out["check"]["serial_iterations"] == out["iterations"]
    assert out["check"]["max_rel_residual_history_diff"] < 1e-10  # RED order only
