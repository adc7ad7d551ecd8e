"""N>1 host logic on CPU (gloo, world_size 2): contiguous env sharding plus the all-reduce of
the additive stats vector must reproduce the single-process statistics, and hence the same
lambda update on every rank (meta_agent.py:32-36).  Per-shard stats come from the oracle here;
on GPUs they come from the accounting kernels (tests/test_gpu_parity.py checks those)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stats_vector(R, modR, C, thr):
    """include/smarl.h smarl_stats_len layout from per-episode products [E, .]."""
    return np.concatenate([C.sum(0), (C > thr[None]).sum(0), R.sum(0), modR.sum(0), [R.shape[0]]]).astype(np.float64)


def make_case():
    from oracle import numpy_oracle as no
    size, A, E, T, gamma = 5, 3, 101, 20, 0.999
    rng = np.random.default_rng(7)
    starts = np.floor(rng.random((E, A, 2)) * size).astype(np.int64)
    actions = rng.integers(0, 5, size=(T, E, A))
    lam, thr = np.array([0.1, 0.2, 0.3]), np.array([8.0, 8.0, 8.0])
    lut = no.coverage_penalty_lut(size, no.coverage_fieldview(size, A))

    def run(lo, hi):
        pos = starts[lo:hi].copy()

        def step_fn(t):
            nonlocal pos
            pos, r, c, _ = no.coverage_discrete_step(pos, actions[t, lo:hi], size, lut, [1.0, 2.0, 3.0])
            return r, c
        out = no.rollout(step_fn, T, gamma, lam)
        return stats_vector(out["R"], out["modR"], out["C"], thr)
    return E, A, lam, thr, run


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from safe_multiagent_rl_b200 import dist as sd
    r, w, _ = sd.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    E, A, lam, thr, run = make_case()
    off, n = sd.shard_range(E, rank, world)
    vec = torch.from_numpy(run(off, off + n))
    sd.allreduce_stats(vec)
    q.put((rank, off, n, vec.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from safe_multiagent_rl_b200.dist import shard_range
    for total in [1, 7, 16, 101, 4194304]:
        for world in [1, 2, 3, 8]:
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(n for _, n in spans) == total
            for (o1, n1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + n1 == o2
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_gloo_allreduce_equals_single_process():
    from oracle import numpy_oracle as no
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    E, A, lam, thr, run = make_case()
    whole = run(0, E)
    K = A
    for rank, off, n, vec in got:
        # integer-valued slots (cost sums, violation counts, episode count) are exact; f64 return sums
        # differ only by summation order
        assert np.array_equal(vec[: 2 * K], whole[: 2 * K]) and vec[-1] == whole[-1] == E
        np.testing.assert_allclose(vec[2 * K:-1], whole[2 * K:-1], rtol=1e-13)
        lam_rank = no.lambda_update(lam, vec[:K] / vec[-1], thr, 0.05)
        lam_one = no.lambda_update(lam, whole[:K] / whole[-1], thr, 0.05)
        assert np.array_equal(lam_rank, lam_one)          # bit-identical lambda on every rank
    assert sorted((off, n) for _, off, n, _ in got) == [(0, 51), (51, 50)]
