"""Multi-GPU partition of HP-2: the i<j pair triangle is cut into one contiguous,
equal-sized pair range per rank; every rank computes its range with no data-path collective, then the
packed (distance, variance) entries are gathered with ONE all-gather (NCCL over NVLink on GPUs, gloo in
the CPU tests) and scattered into the two symmetric N x N matrices."""
import numpy as np


def n_pairs(N):
    return N * (N - 1) // 2


def shard_size(N, world):
    return (n_pairs(N) + world - 1) // world


def pair_shard(N, rank, world):
    """[begin, end) of the linearised upper triangle owned by `rank`."""
    P, per = n_pairs(N), shard_size(N, world)
    return min(P, rank * per), min(P, (rank + 1) * per)


def pair_index(i, j, N):
    return i * N - i * (i + 1) // 2 + (j - i - 1)


def scatter_to_matrices(packed, N):
    """packed: (world, 2, per) gathered shards -> (distances, variances) N x N, symmetric, zero diagonal
    (DistanceMatrix, DistanceFactory.h:12-42)."""
    world, _, per = packed.shape
    P = n_pairs(N)
    d = packed[:, 0, :].reshape(-1)[:P]
    v = packed[:, 1, :].reshape(-1)[:P]
    D = np.zeros((N, N)); V = np.zeros((N, N))
    iu = np.triu_indices(N, 1)
    D[iu] = d; V[iu] = v
    return D + D.T, V + V.T


def all_gather_shards(dist_mod, local, world):
    """local: torch tensor (2, per) on the rank's device -> (world, 2, per) on every rank."""
    import torch
    out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist_mod.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1))
    return out
