"""Multi-GPU plumbing for the level-0 path: one process per GPU, SNP blocks sharded by contiguous
ranges exactly like the reference's own multi-process mode (`--split-l0`, write_l0_master,
src/Data.cpp:268-301: floor(B/njobs) blocks per job, the first B mod njobs jobs take one more).
No data-path collective is needed for level 0; torch.distributed is only used to agree on timings /
status.  Works with the nccl backend on GPUs and with gloo on CPUs (tests)."""
import torch
import torch.distributed as dist


def partition_blocks(n_blocks: int, world: int):
    """[(first_block, n_blocks)] per rank, contiguous, reference rule."""
    if world > n_blocks:
        raise ValueError("number of ranks cannot be greater than number of blocks")
    nall, rem = divmod(n_blocks, world)
    out, start = [], 0
    for r in range(world):
        n = nall + (1 if r < rem else 0)
        out.append((start, n))
        start += n
    return out


def my_blocks(n_blocks: int):
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    first, n = partition_blocks(n_blocks, world)[rank]
    return list(range(first, first + n))


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reported as the max over ranks."""
    if not dist.is_initialized():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def first_error(code: int, device=None) -> int:
    """Smallest non-zero status over ranks (0 = every rank is fine)."""
    if not dist.is_initialized():
        return code
    big = 2 ** 62
    t = torch.tensor([code if code else big], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    v = int(t.item())
    return 0 if v == big else v


def gather_block_owner(n_blocks: int, device=None):
    """all_gather of the block->rank map (every rank must see the same partition)."""
    owner = torch.full((n_blocks,), -1, dtype=torch.int64, device=device)
    for b in my_blocks(n_blocks):
        owner[b] = dist.get_rank() if dist.is_initialized() else 0
    if dist.is_initialized():
        dist.all_reduce(owner, op=dist.ReduceOp.MAX)
    return owner.cpu().tolist()
