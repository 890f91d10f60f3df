"""World-size-2 gloo tests of the sample-sharded calibration plumbing (mobilequant_b200/utils/dist.py): ownership of
samples, the packed range all-reduce (act-range mode) and the gradient all-reduce (LET/LWC/LRL mode).  CPU only."""
import os, socket
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mobilequant_b200.utils import dist as D
    try:
        n = 11
        mine = D.shard_indices(n)
        # act-range mode: every rank folds its own samples, one MAX all-reduce over [-min, max]
        g = torch.Generator().manual_seed(1337)
        samples = torch.randn(n, 7, 64, generator=g)                 # n samples x 7 statistics x 64 values
        packed = torch.stack([torch.stack([samples[mine, s].min(), samples[mine, s].max()]) for s in range(7)])
        red = D.allreduce_ranges(packed)
        # LET/LWC/LRL mode: per-rank gradients of micro-batch `rank`, SUM all-reduce then the batch mean
        params = [torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(3, 1)), torch.nn.Parameter(torch.zeros(()))]
        gg = torch.Generator().manual_seed(100 + rank)
        for p in params:
            p.grad = torch.randn(p.shape, generator=gg)
        params.append(torch.nn.Parameter(torch.zeros(2)))            # a parameter without a gradient is skipped
        D.allreduce_grads(params)
        q.put((rank, mine, red.numpy(), [p.grad.numpy().copy() for p in params[:3]]))
    finally:
        dist.destroy_process_group()


def test_sharded_calibration_exchanges_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # ownership: a partition of the samples
    assert sorted(res[0][1] + res[1][1]) == list(range(11)) and not set(res[0][1]) & set(res[1][1])
    # ranges: identical on both ranks and bit-identical to the single-process pass
    g = torch.Generator().manual_seed(1337)
    samples = torch.randn(11, 7, 64, generator=g)
    ref = torch.stack([torch.stack([samples[:, s].min(), samples[:, s].max()]) for s in range(7)]).numpy()
    assert np.array_equal(res[0][2], ref) and np.array_equal(res[1][2], ref)
    # gradients: the mean over ranks, identical on both
    want = []
    for shape in [(5,), (3, 1), ()]:
        want.append(None)
    per_rank = []
    for r in range(world):
        gg = torch.Generator().manual_seed(100 + r)
        per_rank.append([torch.randn(s, generator=gg) for s in [(5,), (3, 1), ()]])
    for i in range(3):
        mean = ((per_rank[0][i] + per_rank[1][i]) / 2).numpy()
        assert np.allclose(res[0][3][i], mean, rtol=0, atol=1e-7) and np.array_equal(res[0][3][i], res[1][3][i])


def test_shard_indices_single_process():
    from mobilequant_b200.utils.dist import shard_indices
    assert shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    assert shard_indices(5, 1, 2) == [1, 3]
    assert shard_indices(0, 0, 4) == []
