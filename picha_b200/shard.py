"""Sharding of a batch of independent images across ranks / GPUs: contiguous
blocks, no exchange.  Rank g of G owns images [g*N/G, (g+1)*N/G)."""


def shard_range(n, rank, world):
    if world <= 0 or rank < 0 or rank >= world:
        raise ValueError("bad rank/world")
    return (n * rank) // world, (n * (rank + 1)) // world
