"""The N > 1 path on CPU: world_size-2 `gloo` process group, the product's sharding plumbing
(hpg_variant_b200/sharding.py: contiguous combination-index ranges + one all-gather of the per-rank
top-N records).  The per-rank search and the final merge are CUDA kernels in the product; here the
oracle (test infrastructure) stands in for both so that the host-side logic can run without a GPU:
per-rank oracle search over the rank's range -> gloo all-gather -> canonical merge == full search."""
import os
import socket
import sys

import numpy as np
import pytest

from hpg_variant_b200 import sharding
from hpg_variant_b200._lib import MODEL_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("total", [0, 1, 7, 49995000, 20820835000, (1 << 63) + 12345])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_ranges_tile_the_index_space(total, world):
    cuts = [sharding.shard_range(total, r, world) for r in range(world)]
    assert cuts[0][0] == 0 and cuts[-1][1] == total
    for (a, b), (c, d) in zip(cuts[:-1], cuts[1:]):
        assert b == c and a <= b
    sizes = [b - a for a, b in cuts]
    assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(total, world, world)


def canonical_merge(lists, rank_size):
    """[world, F, N] records -> [F, N]: accuracy descending, SNP tuple ascending, empty slots last
    (the order hpgv_epi.h promises; merge_kernel implements it on the GPU)."""
    world, F, N = lists.shape
    out = np.zeros((F, rank_size), MODEL_DTYPE)
    out["accuracy"] = np.nan
    out["snp"] = -1
    for f in range(F):
        rows = [r for r in lists[:, f].reshape(-1) if r["snp"][0] >= 0]
        rows.sort(key=lambda r: (-(r["accuracy"] if not np.isnan(r["accuracy"]) else -np.inf), tuple(int(x) for x in r["snp"])))
        for i, r in enumerate(rows[:rank_size]):
            out[f, i] = r
    return out


def _worker(rank, world, port, order, nv, A, U, F, rank_size, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from hpg_variant_b200 import sharding as sh, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = oracle_lib.Checker("oracle")
        g = synth.make_dataset(nv, A, U, seed=31 + order, order=order, missing=0.01, planted=1)
        fos = (np.concatenate([np.arange(A), np.arange(U)]) % F).astype(np.int32)
        total = int(oracle._fn("num_combinations")(nv, order))
        first, last = sh.shard_range(total, rank, world)
        part, _ = oracle.search(g, A, U, order, fos, 1, rank_size, first=first, last=last, threads=1, num_folds=F)
        local = torch.from_numpy(part.view(np.uint8).reshape(-1).copy())
        gathered = sh.all_gather_models(dist, local, world)
        assert gathered.numel() == world * F * rank_size * sh.RECORD_BYTES
        lists = sh.lists_view(gathered.numpy(), world, F, rank_size)
        # rank r's block of the gathered buffer is rank r's list, on every rank
        assert lists[rank].tobytes() == part.tobytes()
        q.put((rank, first, last, gathered.numpy().tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("order,nv,A,U,F,rank_size", [(2, 40, 60, 70, 5, 12), (3, 14, 48, 48, 3, 20)])
def test_two_rank_gather_and_merge_equals_full_search(oracle, order, nv, A, U, F, rank_size):
    import torch.multiprocessing as mp
    from hpg_variant_b200 import synth
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, order, nv, A, U, F, rank_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # both ranks hold the same gathered buffer and their ranges tile the space
    assert got[0][3] == got[1][3]
    assert got[0][1] == 0 and got[0][2] == got[1][1]
    lists = sharding.lists_view(np.frombuffer(got[0][3], np.uint8), world, F, rank_size)
    merged = canonical_merge(lists, rank_size)
    g = synth.make_dataset(nv, A, U, seed=31 + order, order=order, missing=0.01, planted=1)
    fos = (np.concatenate([np.arange(A), np.arange(U)]) % F).astype(np.int32)
    full, _ = oracle.search(g, A, U, order, fos, 1, rank_size, threads=2, num_folds=F)
    assert got[1][2] == int(oracle._fn("num_combinations")(nv, order))
    assert np.array_equal(merged["snp"][..., :order], full["snp"][..., :order])
    assert np.array_equal(merged["risky_mask"], full["risky_mask"])
    assert np.array_equal(merged["conf"], full["conf"])
    assert np.array_equal(merged["accuracy"], full["ba"], equal_nan=True)


class _CpuEngine:
    """Stand-in for EpistasisEngine in the gloo test of ShardedSearch.run_from_host: same calls, the oracle does the search
    and a numpy merge does the merge, on the CPU buffers whose addresses it is handed (test infrastructure only)."""

    def __init__(self, oracle, F, N):
        self.oracle, self.F, self.N = oracle, F, N
        self.seen = None

    @staticmethod
    def _view(ptr, nbytes):
        import ctypes
        return np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr))

    def load_dataset_device(self, ptr, nv, A, U):
        self.g = self._view(ptr, nv * (A + U)).reshape(nv, A + U).copy()
        self.A, self.U = A, U

    def set_folds(self, F, fos):
        self.fos = np.asarray(fos, np.int32)

    def search_device(self, order, subset, N, first, last, out_ptr):
        self.order = order
        part, _ = self.oracle.search(self.g, self.A, self.U, order, self.fos, subset, N, first=first, last=last, threads=1, num_folds=self.F)
        self._view(out_ptr, part.nbytes)[:] = part.view(np.uint8).reshape(-1)

    def merge_device(self, order, subset, world, N, lists_ptr, out_ptr):
        lists = self._view(lists_ptr, world * self.F * N * 40).view(MODEL_DTYPE).reshape(world, self.F, N)
        self._view(out_ptr, self.F * N * 40)[:] = canonical_merge(lists, N).view(np.uint8).reshape(-1)


def _worker_from_host(rank, world, port, nv, A, U, F, rank_size, q):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    import oracle_lib
    from hpg_variant_b200 import sharding as sh, synth

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = oracle_lib.Checker("oracle")
        g = synth.make_dataset(nv, A, U, seed=77, missing=0.01, planted=1)
        fos = (np.concatenate([np.arange(A), np.arange(U)]) % F).astype(np.int32)
        eng = _CpuEngine(oracle, F, rank_size)
        shard = sh.ShardedSearch(eng, dist, rank, world, F, rank_size, "cpu")
        total = int(oracle._fn("num_combinations")(nv, 2))
        res = shard.run_from_host(torch.from_numpy(g), A, U, F, fos, 2, 1, total)
        assert np.array_equal(eng.g, g)           # the all-gathered slices are the whole matrix, on every rank
        q.put((rank, res.tobytes()))
    finally:
        dist.destroy_process_group()


def test_two_rank_run_from_host_uploads_slices_and_gathers(oracle):
    """ShardedSearch.run_from_host: every rank contributes 1/world of the SNP rows (41 rows over 2 ranks: uneven), the
    all-gather rebuilds the matrix, and the final ranking equals the full search on both ranks."""
    import torch.multiprocessing as mp
    from hpg_variant_b200 import synth
    nv, A, U, F, rank_size, world, port = 41, 60, 70, 4, 10, 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_from_host, args=(r, world, port, nv, A, U, F, rank_size, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == got[1][1]
    g = synth.make_dataset(nv, A, U, seed=77, missing=0.01, planted=1)
    fos = (np.concatenate([np.arange(A), np.arange(U)]) % F).astype(np.int32)
    full, _ = oracle.search(g, A, U, 2, fos, 1, rank_size, threads=2, num_folds=F)
    merged = np.frombuffer(got[0][1], MODEL_DTYPE).reshape(F, rank_size)
    assert np.array_equal(merged["snp"][..., :2], full["snp"][..., :2]) and np.array_equal(merged["conf"], full["conf"])
    assert np.array_equal(merged["accuracy"], full["ba"], equal_nan=True)
