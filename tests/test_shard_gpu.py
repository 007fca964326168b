"""GPU tier: batch sharding.  Graphs are independent, so a shard's results equal the same graphs' results in the unsharded
batch -- except for EuclideanEdge, whose distance averages over the current observation of EVERY graph of the batch
(reference edge_selectors/distance.py:48-49): with DenseGCM.batch_group set, every step all-gathers the ranks'
observations first and the sharded run reproduces the unsharded reference bit for bit in its edge set."""
import os
import sys

import pytest
import torch

import gcm_oracle as oracle
from helpers import make_dense_gnn, make_selector, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _clustered(gen, T, B, F, K=6, noise=0.03):
    centres = torch.randn(K, F, generator=gen) * 1.5
    sched = torch.randint(0, K, (T,), generator=gen)
    return (centres[sched].unsqueeze(1).expand(T, B, F) + noise * torch.randn(T, B, F, generator=gen)).contiguous()


def test_euclidean_shards_with_gathered_observations_equal_the_unsharded_batch():
    """Two 'ranks' in one process (each a DenseGCM on half of the batch, the all-gather replaced by a concatenation of the
    two halves): beliefs and final states of the halves == the unsharded oracle; without the gather they differ."""
    from gcm.gcm import DenseGCM

    dev = torch.device("cuda:0")
    B, N, F, H, T = 12, 10, 16, 16, 25
    spec = [("euclidean", 2.0)]
    gen = torch.Generator().manual_seed(77)
    obs = _clustered(gen, T, B, F)
    # on odd steps the first half's observations sit far away: the batch-mean distance of the second half's nodes then
    # exceeds the threshold (no edge), while a shard that only saw its own observations would still link them
    obs[1::2, : B // 2] += 6.0 / F ** 0.5
    p = oracle.make_params(F, H)
    mods = []
    for _ in range(2):
        gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
        mods.append(DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N))
    halves = [slice(0, B // 2), slice(B // 2, B)]
    hid, o_hidden = [None, None], None
    cur_all = [None]
    for m in mods:
        m.batch_group = True
        m._all_current_obs = lambda x: cur_all[0]                    # what gather_current_obs returns on every rank
    with torch.no_grad():
        for t in range(T):
            cur_all[0] = obs[t].to(dev)
            ref, o_hidden = oracle.dense_gcm_step(obs[t], o_hidden, spec, p, graph_size=N)
            for r in range(2):
                belief, hid[r] = mods[r](obs[t, halves[r]].to(dev), hid[r])
                assert rel_err(belief, ref[halves[r]]) < 2e-5, (t, r)
    for r in range(2):
        nodes, adj, _, nn = hid[r]
        assert torch.equal(adj.cpu(), o_hidden[1][halves[r]]) and torch.equal(nodes.cpu(), o_hidden[0][halves[r]])
    assert float(o_hidden[1].sum()) > 0
    # sanity: a shard that only sees its own observations computes different mean distances
    gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
    lone = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
    h = None
    with torch.no_grad():
        for t in range(T):
            _, h = lone(obs[t, halves[1]].to(dev), h)
    assert not torch.equal(tuple(h)[1].cpu(), o_hidden[1][halves[1]])


def _rank(rank, world, port, q):
    for pth in (os.path.join(ROOT, "graph-conv-memory_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, pth)
    import torch.distributed as dist

    from gcm import dist as gdist
    from gcm.gcm import DenseGCM

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        B, N, F, H, T = 300, 10, 16, 16, 14                         # 150 graphs per rank: the tensor-core distance kernel
        spec = [("euclidean", 2.0)]
        gen = torch.Generator().manual_seed(78)
        obs = _clustered(gen, T, B, F)
        p = oracle.make_params(F, H)
        gnn, _ = make_dense_gnn(F, H, p, ("tanh", "tanh"))
        mod = DenseGCM(gnn.to(dev), edge_selectors=make_selector(spec), graph_size=N)
        mod.batch_group = True
        lo, hi = gdist.shard_bounds(B, rank, world)
        hidden, o_hidden = None, None
        with torch.no_grad():
            for t in range(T):
                belief, hidden = mod(obs[t, lo:hi].to(dev), hidden)
                ref, o_hidden = oracle.dense_gcm_step(obs[t], o_hidden, spec, p, graph_size=N)
                assert rel_err(belief, ref[lo:hi]) < 2e-5, t
        assert torch.equal(tuple(hidden)[1].cpu(), o_hidden[1][lo:hi])
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.timeout(300)
def test_euclidean_batch_sharding_over_nccl_two_ranks():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_rank, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for pr in procs:
        pr.join(30)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
