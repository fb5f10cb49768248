"""Multi-GPU path through NCCL (needs >= 2 GPUs; skipped otherwise): sample routing from one rank, sharded splat, and both
final-assembly paths (resolve + ncclAllGather, and the fused resolve-with-peer-stores kernel) against the single film."""
import json
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _ngpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,res,spp,filt", [(2, (320, 203), 4, "gaussian"), (2, (256, 130), 16, "lanczos"),
                                                (4, (300, 210), 4, "gaussian"), (8, (200, 164), 4, "gaussian")])
def test_routed_sharded_render_equals_single_film(gpu, tmp_path, world, res, spp, filt):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "result.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "dist_worker_gpu.py"), str(out), str(res[0]), str(res[1]),
           str(spp), filt]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    for rank in range(world):
        d = json.loads(Path(f"{out}.{rank}").read_text())
        assert d["frames_identical"] and d["all_ranks_same_frame"] and d["shape"] == [res[1], res[0], 3], d
        if rank == 0:
            assert d["equals_single_film"] and d["nonzero"], d
