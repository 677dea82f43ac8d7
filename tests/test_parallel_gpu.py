"""Ulysses sequence parallelism against the single-GPU forward: the fused peer-memory exchange (ug_qkv_scatter,
ug_attention_bf16_peer, ug_peer_barrier, ug_peer_bcast_rows over CUDA IPC pools) runs with world size 1 on any box — every
kernel of the exchange executes, all peer pointers map to the local pool — and with 2 ranks when the box has 2 GPUs."""
import json
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _run(world, port, *extra, tool="sp_check.py"):
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(ROOT / "tools" / tool), "--workload", "tiny", *extra],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])


def test_peer_exchange_world1_equals_single_gpu_forward_and_graph_replay():
    rec = _run(1, 29540, "--exchange", "peer", "--graph")
    assert rec["ok"] and rec["rel_l2"] == 0.0 and rec["graph_equals_eager"] and rec["peer_barrier_timeouts"] == 0, rec


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_ulysses_sp_equals_single_gpu_forward(exchange):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rec = _run(2, 29541, "--exchange", exchange, *(["--graph"] if exchange == "peer" else []))
    assert rec["ok"], rec


def test_pvariant_segment_sharded_exchange_world1_equals_single_gpu_forward():
    """P-variant with the segment-sharded peer exchange at world size 1: scatter per segment, masked attention with the
    segment-sharded output mapping, barriers, velocity gather and CUDA-graph replay — bit-identical to UniCombineFlux."""
    rec = _run(1, 29542, "--graph", tool="sp_check_pvariant.py")
    assert rec["ok"] and rec["rel_l2"] == 0.0 and rec["graph_equals_eager"], rec


def test_pvariant_segment_sharded_ulysses_2gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    rec = _run(2, 29543, "--graph", tool="sp_check_pvariant.py")
    assert rec["ok"] and rec["rel_l2"] == 0.0, rec
