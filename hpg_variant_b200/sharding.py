"""Multi-GPU sharding of the exhaustive search: one process per GPU, contiguous combination-index
ranges, ONE all-gather of the per-rank top-N lists, the same deterministic merge on every rank.

Replaces the MPI work distribution and tree merge of the reference
(mpi/epistasis_runner.c:129-157 MPI_Scatterv of block coordinates, :410-452 MPI_Send/Recv merge of
the per-fold heaps).  The combination space shards naturally: the only cross-rank
state is F x N fixed-size records (40 bytes each), so nothing but this gather crosses NVLink.

`torch.distributed` is plumbing here (NCCL on the GPUs, gloo in the CPU tests); the search and the
merge are CUDA kernels behind the C-ABI (hpgv_epi_search_device / hpgv_epi_merge_device).
"""
import numpy as np

from ._lib import MODEL_DTYPE

RECORD_BYTES = MODEL_DTYPE.itemsize


def shard_range(total, rank, world):
    """[first, last) of rank's contiguous slice of the linear combination index space [0, total).
    Slices differ by at most one combination and tile the space exactly (lexicographic tuple order,
    pair (i, j) <-> i*(2n-i-1)/2 + (j-i-1); hpgv_epi.h documents the numbering)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank must be in [0, world)")
    return total * rank // world, total * (rank + 1) // world


def all_gather_models(dist, local, world, out=None, group=None):
    """local: uint8 tensor [F * N * 40] (this rank's per-fold lists as hpgv_epi_model_t records) on
    the backend's device.  Returns the uint8 tensor [world * F * N * 40], rank-major: exactly the
    d_lists argument of hpgv_epi_merge_device."""
    import torch
    if local.dtype != torch.uint8 or local.dim() != 1 or local.numel() % RECORD_BYTES:
        raise ValueError("local must be a flat uint8 tensor of whole 40-byte records")
    if out is None:
        out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    if world == 1:
        out.copy_(local)
    else:
        dist.all_gather_into_tensor(out, local, group=group)
    return out


def lists_view(buf, world, num_folds, rank_size):
    """numpy structured view [world, F, N] of a gathered host buffer (tests, reports)."""
    a = np.frombuffer(bytes(buf), dtype=MODEL_DTYPE) if not isinstance(buf, np.ndarray) else buf.view(MODEL_DTYPE)
    return a.reshape(world, num_folds, rank_size)


class ShardedSearch:
    """One rank's share of a search over `world` GPUs.  Buffers are allocated once; `run` enqueues
    search -> all-gather -> merge on the engine's stream / the process group and leaves the final
    F x N ranking (identical on every rank) in `self.d_final`."""

    def __init__(self, engine, dist, rank, world, num_folds, rank_size, device):
        import torch
        self.eng, self.dist, self.rank, self.world = engine, dist, rank, world
        self.F, self.N = num_folds, rank_size
        nbytes = num_folds * rank_size * RECORD_BYTES
        self.d_local = torch.zeros(nbytes, dtype=torch.uint8, device=device)
        self.d_all = torch.zeros(world * nbytes, dtype=torch.uint8, device=device)
        self.d_final = torch.zeros(nbytes, dtype=torch.uint8, device=device)

    def run(self, order, eval_subset, total):
        first, last = shard_range(total, self.rank, self.world)
        if self.world == 1:
            self.eng.search_device(order, eval_subset, self.N, first, last, self.d_final.data_ptr())
            return self.d_final
        self.eng.search_device(order, eval_subset, self.N, first, last, self.d_local.data_ptr())
        all_gather_models(self.dist, self.d_local, self.world, out=self.d_all)
        self.eng.merge_device(order, eval_subset, self.world, self.N, self.d_all.data_ptr(), self.d_final.data_ptr())
        return self.d_final

    def run_from_host(self, g_pinned, num_affected, num_unaffected, num_folds, fold_of_sample, order, eval_subset, total):
        """The whole path from HOST buffers: g_pinned is the pinned uint8 tensor [num_variants, A + U] every rank holds.
        Each rank uploads only its 1/world slice of the SNP rows over its own PCIe link and the slices are all-gathered
        over NVLink (every rank needs every SNP: a pair joins two arbitrary rows), then pack, search, gather, merge.
        Returns the final ranking as a host array (identical on every rank)."""
        import torch
        nv, S = g_pinned.shape
        per = -(-nv // self.world)
        if getattr(self, "_raw_shape", None) != (nv, S):
            self.d_raw_full = torch.empty((self.world * per, S), dtype=torch.uint8, device=self.d_local.device)
            self.d_raw_slice = torch.empty((per, S), dtype=torch.uint8, device=self.d_local.device)
            self._raw_shape = (nv, S)
        lo, hi = min(nv, self.rank * per), min(nv, (self.rank + 1) * per)
        if self.world == 1:
            self.d_raw_full[:nv].copy_(g_pinned, non_blocking=True)
        else:
            self.d_raw_slice[: hi - lo].copy_(g_pinned[lo:hi], non_blocking=True)
            self.dist.all_gather_into_tensor(self.d_raw_full.view(-1), self.d_raw_slice.view(-1))
        self.eng.load_dataset_device(self.d_raw_full.data_ptr(), nv, num_affected, num_unaffected)
        self.eng.set_folds(num_folds, fold_of_sample)
        self.run(order, eval_subset, total)
        return self.result()

    def result(self):
        """Host copy of the final ranking as a structured array [F, N]."""
        return self.d_final.cpu().numpy().view(MODEL_DTYPE).reshape(self.F, self.N)
