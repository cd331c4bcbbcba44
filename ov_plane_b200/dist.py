"""Multi-GPU sharding of ONE large MSCKF point update: features are dealt round-robin to the ranks, every rank
compresses its own rows against the replicated state into an (n+1)x(n+1) factor block [R^T ; z^T] in an agreed column order,
ONE all-gather moves the blocks (NCCL over NVLink on GPUs, gloo in the CPU tests), and every rank applies the identical
second-level compression + EKF update (replicated, so no broadcast of P is needed).

torch.distributed is plumbing only; the compute callbacks come from the caller (the CUDA library on GPUs; the CPU tests
inject the oracle to validate the sharding algebra and the gather layout)."""


def shard_indices(n_features, rank, world):
    """feature i -> rank i mod world"""
    return list(range(rank, n_features, world))


def all_gather_blocks(block, world):
    """block: 1-D tensor of (n+1)^2 doubles on this rank -> 1-D tensor of world*(n+1)^2 doubles, rank-major"""
    import torch
    import torch.distributed as dist
    if world == 1:
        return block.clone()
    out = torch.empty(world * block.numel(), dtype=block.dtype, device=block.device)
    dist.all_gather_into_tensor(out, block)
    return out


def sharded_update(compress_fn, update_fn, n_features, rank, world):
    """compress_fn(feature_indices) -> 1-D tensor block ; update_fn(all_blocks_tensor, world) -> None"""
    mine = shard_indices(n_features, rank, world)
    blk = compress_fn(mine)
    allb = all_gather_blocks(blk, world)
    update_fn(allb, world)
    return mine
