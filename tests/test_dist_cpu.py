"""CPU tier, world_size 2 over gloo: the host-side logic of the batch-sharded multi-GPU path."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "graph-conv-memory_b200"))
    from gcm import dist as gdist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gen = torch.Generator().manual_seed(5)
        full = torch.randn(11, 4, generator=gen)                      # global batch of 11 graphs
        mine = gdist.shard(full, rank, world)
        # 1. the shards tile the batch
        sizes = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.shape[0]]))
        assert sum(int(s) for s in sizes) == 11
        # 2. EuclideanEdge's cross-batch term: every rank sees all current observations, in batch order
        allx = gdist.gather_current_obs(mine)
        assert torch.equal(allx, full)
        assert gdist.shard_sizes(mine.shape[0], mine.device) == [6, 5]
        assert torch.equal(gdist.gather_current_obs(mine, None, [6, 5]), full)      # cached sizes: no size exchange
        even = gdist.shard(full[:10], rank, world)                                   # equal shards: one all_gather_into_tensor
        assert torch.equal(gdist.gather_current_obs(even, None, [5, 5]), full[:10])
        # 3. one flattened all-reduce of the weight gradients == gradient of the unsharded loss
        lin = torch.nn.Linear(4, 3)
        with torch.no_grad():
            lin.weight.copy_(torch.arange(12.0).view(3, 4) / 10)
            lin.bias.zero_()
        lin(mine).pow(2).sum().backward()
        n = gdist.allreduce_grads(lin.parameters())
        ref = torch.nn.Linear(4, 3)
        ref.load_state_dict(lin.state_dict())
        ref(full).pow(2).sum().backward()
        assert n == 15
        assert torch.allclose(lin.weight.grad, ref.weight.grad, atol=1e-5)
        assert torch.allclose(lin.bias.grad, ref.bias.grad, atol=1e-5)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_batch_sharding_and_grad_allreduce_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(world)]
    for p in procs:
        p.join(20)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
