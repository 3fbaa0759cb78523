"""Real-rank parity of the three sharded paths: one process per GPU over NCCL / NVLink peer memory,
min(8, visible GPUs) ranks - the sequence-parallel DiT forward (NCCL all-to-all form AND the peer-memory form) equals
the single-GPU forward bit for bit, the row-sharded VAE equals the single-GPU VAE bit for bit, the channel-sharded FLF
scoring equals single-rank scoring.  The check itself is tools/ulysses_check.py (also run by hand through gpurun;
its 8-rank log is committed under profiles/).  Skipped where only one GPU is visible; the host-side logic of the same
paths runs on 2 and 4 CPU ranks over gloo in test_ulysses_gloo.py / test_vae_rows.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_paths_on_real_ranks(cuda):
    n = min(8, torch.cuda.device_count())
    if n < 2:
        pytest.skip("one GPU visible: the real-rank check needs >= 2 (run under gpurun --gpus N)")
    n = 1 << (n.bit_length() - 1)                      # 2, 4 or 8 ranks
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "ulysses_check.py")],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    for what in ("sequence-parallel forward equal to single-GPU: True", "peer-memory forward equal to single-GPU: True",
                 "cfg-parallel forwards equal to single-GPU: True",
                 "row-sharded VAE equal to single-GPU: encode True decode True", "rank-sharded FLF scores equal: True"):
        assert out.count(what) == n, (what, out[-4000:])
    assert out.count("matches single-GPU: True") == n, out[-4000:]          # LongCat 2-D context parallel
