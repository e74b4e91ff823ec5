"""N>1 host logic on CPU: world_size-2 gloo processes agree on the block partition, the
max-over-ranks timing reduction and error propagation (the reference pins sharded == unsharded,
test/test_bash.sh:91-137)."""
import os
import socket

import torch.multiprocessing as mp

from regenie_b200 import sharding


def test_partition_matches_reference_rule():
    assert sharding.partition_blocks(10, 4) == [(0, 3), (3, 3), (6, 2), (8, 2)]   # nall=2, remainder=2
    assert sharding.partition_blocks(50, 8)[0] == (0, 7) and sharding.partition_blocks(50, 8)[-1] == (44, 6)
    parts = sharding.partition_blocks(500, 8)
    assert sum(n for _, n in parts) == 500 and all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(7))


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = sharding.my_blocks(11)
    owner = sharding.gather_block_owner(11)
    tmax = sharding.max_over_ranks(10.0 + rank)
    err = sharding.first_error(0 if rank == 0 else 4242)
    rows = sharding.step2_distributed(11, lambda b: ["block %d row %d (rank %d)" % (b, k, rank) for k in range(b % 3 + 1)])
    q.put((rank, mine, owner, tmax, err, rows))
    dist.destroy_process_group()


def test_two_rank_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res[0][1] == list(range(0, 6)) and res[1][1] == list(range(6, 11))
    assert res[0][2] == res[1][2] == [0] * 6 + [1] * 5
    assert res[0][3] == res[1][3] == 11.0
    assert res[0][4] == res[1][4] == 4242
    # Step 2: rank 0 holds the rows of every block in file order, produced by the rank that owns the block
    assert res[1][5] is None
    want = [["block %d row %d (rank %d)" % (b, k, 0 if b < 6 else 1) for k in range(b % 3 + 1)] for b in range(11)]
    assert res[0][5] == want
