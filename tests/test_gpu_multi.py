"""Multi-GPU parity, on a box with at least two GPUs: one process per GPU under
torchrun, the library's own NCCL communicator; every rank must write files byte-identical to the
reference fixtures, and the sharded engine must equal the unsharded one bit for bit."""
import json
import os
import socket
import subprocess
import sys

import pytest

import parity_util as pu

pytestmark = pytest.mark.gpu
ROOT = pu.ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_run_matches_reference_fixtures(world, tmp_path):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from multi_worker import CASES
    golden = pu.load_golden()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multi_worker.py"), str(tmp_path)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    for rank in range(world):
        with open(tmp_path / f"result{rank}.json") as f:
            r = json.load(f)
        assert r["world"] == world
        assert r["engine"]["bits_equal"] and r["engine"]["pop_equal"] and r["engine"]["dec_equal"]
        assert r["engine"]["selected"] > 0
        for name in CASES:
            g = golden[name]
            assert r["cases"][name]["outputs"] == g["outputs"], (rank, name)
            stats = dict((k, v) for k, v in g["stats"] if k in ("m_filterSize:", "num_passed_reads:"))
            assert r["cases"][name]["filter_bits"] == stats["m_filterSize:"]
            assert r["cases"][name]["num_passed_reads"] == stats["num_passed_reads:"]
        # grb_run_two_stage with the ingest sharded over the ranks: the reference's files
        assert r["two_stage"]["silver"] == golden["silver_default"]["outputs"]
        assert r["two_stage"]["golden"] == golden["golden_default"]["outputs"]
        # slice mode: each rank saw only its own records of the input
        sl = r["slice"]
        assert 0 < sl["bytes"] and sl["digest"] == sl["expect"] and sl["selected"] == sl["expect_selected"]
        assert sl["two_silver"] == sl["expect"] and sl["two_golden"] == sl["expect_golden"]
        assert sl["golden_reads"] == sl["expect_golden_reads"]
