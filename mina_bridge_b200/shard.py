"""Sharding of a proof batch across ranks and the accept-byte reduction.

Proofs are independent (the operator verifies one goroutine per proof and ANDs the results,
AL/operator/pkg/operator.go:448-465), so the batch is partitioned with no data-path exchange:
proof i belongs to rank i mod W.  The only collective is one all-reduce(MIN) over the N result bytes:
every rank writes its own entries and leaves the others at 1, so after MIN each entry holds its owner's
bit, and min(result) is the operator's AND.
"""
from __future__ import annotations


def shard_indices(n: int, rank: int, world: int) -> list[int]:
    return list(range(rank, n, world))


def merge_result_bytes(torch, dist, result, idx, bits, world: int):
    """result: uint8 tensor [N] (any device); idx: int64 tensor of this rank's indices; bits: uint8 tensor."""
    result.fill_(1)
    result[idx] = bits
    if world > 1:
        dist.all_reduce(result, op=dist.ReduceOp.MIN)
    return result
