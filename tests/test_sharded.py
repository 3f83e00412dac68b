"""N>1 path on CPU: world_size-2 gloo run of the batch-shard + logits all-gather driver."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

import bnn_b200
from bnn_b200 import sharded


def test_shard_bounds_cover_the_batch():
    for total in (1, 7, 8, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    model = nn.Sequential(nn.Flatten(), nn.Linear(12, 5)).eval()
    x = torch.randn(total, 3, 2, 2)
    engine = sharded.ShardedInference(model)
    full = engine(x)
    with torch.no_grad():
        want = model(x)
    ok = torch.allclose(full, want, atol=0, rtol=0) and full.shape == want.shape
    torch.save(torch.tensor(int(ok)), os.path.join(out_dir, f"ok{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_world_size_two_gloo(tmp_path, total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert int(torch.load(os.path.join(str(tmp_path), f"ok{r}.pt"))) == 1
