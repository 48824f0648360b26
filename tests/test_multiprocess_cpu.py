"""CPU tier, world size 2 over gloo: the sharding + accept-byte reduction the multi-GPU path uses
(mina_bridge_b200/shard.py; mirrors AL/operator/pkg/operator.go:448-465).  The per-proof bit here is a
host-only stage (does the C++ decoder accept the proof?), so no device is needed."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, proofs, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import mina_bridge_b200 as mb
    from mina_bridge_b200 import shard

    mine = shard.shard_indices(len(proofs), rank, world)
    bits = torch.tensor([int(mb.host_decode(0, proofs[i]) is not None) for i in mine], dtype=torch.uint8)
    result = torch.empty(len(proofs), dtype=torch.uint8)
    shard.merge_result_bytes(torch, dist, result, torch.tensor(mine, dtype=torch.int64), bits, world)
    if rank == 0:
        torch.save(result, out_path)
    else:
        # every rank must end up with the same vector
        ref = result.clone()
        dist.broadcast(ref, src=0)
        assert torch.equal(ref, result)
    if rank == 0:
        ref = result.clone()
        dist.broadcast(ref, src=0)
    dist.destroy_process_group()


def test_shard_and_allreduce_min_world_size_2(native, tmp_path):
    proof = golden("mina_state.proof")
    proofs, want = [], []
    for i in range(37):  # ragged: 19 + 18
        if i % 5 == 3:
            proofs.append(proof[: 1000 + i])  # truncated -> decoder rejects
            want.append(0)
        else:
            proofs.append(proof)
            want.append(1)
    out = str(tmp_path / "result.pt")
    mp.spawn(_worker, args=(2, _free_port(), proofs, out), nprocs=2, join=True)
    assert torch.load(out).tolist() == want


def test_shard_indices_cover_the_batch_exactly_once():
    from mina_bridge_b200 import shard

    for n in (0, 1, 7, 1024):
        for w in (1, 2, 4, 8):
            seen = sorted(i for r in range(w) for i in shard.shard_indices(n, r, w))
            assert seen == list(range(n))
