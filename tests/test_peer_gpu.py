"""The one-shot peer-memory all-reduce (csrc/peer.cuh) through the C-ABI.

On one GPU the exchange degenerates to world = 1 (the kernel pushes into and polls its own buffer),
which still exercises the buffer layout, the step counter on the device, the two alternating slots and
CUDA-graph replay; with two or more GPUs the multi-process check tool is run under torchrun."""
import ctypes
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _alloc(lib, check, n):
    buf = ctypes.c_void_p()
    handle = ctypes.create_string_buffer(64)
    check(lib.doper_peer_alloc(n, ctypes.byref(buf), handle), "doper_peer_alloc")
    return buf


@pytest.mark.parametrize("n", [1, 5, 6603, 40000])
def test_single_rank_identity_and_graph_replay(n):
    from doper_b200._abi import check, lib
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    buf = _alloc(lib, check, n)
    ptrs = (ctypes.c_void_p * 1)(buf.value)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    out = torch.empty(n, dtype=torch.float32, device=dev)
    try:
        for step in range(5):  # both slots, several values of the device-side step counter
            x = torch.randn(n, device=dev)
            check(lib.doper_allreduce_oneshot(torch.cuda.current_stream().cuda_stream, 0, 1, ptrs, n, x.data_ptr(),
                                              out.data_ptr(), status.data_ptr(), 5.0), "doper_allreduce_oneshot")
            assert torch.equal(out, x), step
        # in place, replayed from a CUDA graph
        x = torch.randn(n, device=dev)
        y = x.clone()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            check(lib.doper_allreduce_oneshot(torch.cuda.current_stream().cuda_stream, 0, 1, ptrs, n, y.data_ptr(),
                                              y.data_ptr(), status.data_ptr(), 5.0), "doper_allreduce_oneshot")
        for _ in range(3):
            y.copy_(x)
            graph.replay()
            assert torch.equal(y, x)
        assert int(status.item()) == 0
    finally:
        torch.cuda.synchronize()
        lib.doper_peer_free(buf)


def test_argument_validation():
    from doper_b200._abi import lib
    ptrs = (ctypes.c_void_p * 1)(None)
    assert lib.doper_allreduce_oneshot(None, 0, 1, ptrs, 8, 16, 16, None, 0.0) != 0   # null peer buffer
    assert lib.doper_allreduce_oneshot(None, 0, 17, ptrs, 8, 16, 16, None, 0.0) != 0  # more than 16 ranks
    assert lib.doper_allreduce_oneshot(None, 1, 1, ptrs, 8, 16, 16, None, 0.0) != 0   # rank outside the group
    assert lib.doper_peer_buffer_bytes(0) == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_two_ranks_match_nccl():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", str(ROOT / "tools/peer_allreduce_check.py")],
                         capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["values_match_nccl"] and line["same_bits_on_every_rank"] and line["graph_replay_ok"]
    assert line["status"] == 0


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one node")
def test_data_parallel_iteration_graph_with_peer_exchange():
    """tools/trainer_multi_gpu_check.py: the 8-action iteration as one CUDA graph per rank with the
    peer-memory exchange captured inside == the eager loop with NCCL == one GPU with all the balls."""
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29549",
                          str(ROOT / "tools/trainer_multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, MASTER_ADDR="127.0.0.1"), cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["graph_equals_eager_nccl"] and line["same_bits_on_every_rank"] and line["peer_status"] == 0
    assert line["first_iteration_grad_vs_single_gpu_big_batch_rel_err"] < 1e-4  # fp32 sums in a different order
    assert line["loss_rel_err"] < 1e-5
