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


# ---------------------------------------------------------------------------------------- Step 1 across GPUs
def phenotype_owner(n_pheno: int, world: int):
    """Level 1 is sharded by phenotype: phenotype p is fitted by rank p mod world."""
    return [p % world for p in range(n_pheno)]


def _sum_to_all(arr, device):
    """Element-wise sum over ranks of a host array (entries not owned by a rank are exactly 0, so this is a
    gather).  CPU tensors under gloo, CUDA tensors under nccl."""
    import numpy as np
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == "nccl":
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def attach_peers(st):
    """Give every phenotype's level-0 predictors a home: phenotype p lives in the HBM of rank p mod world; this rank
    maps the other ranks' W allocations (CUDA IPC over NVLink / NVSwitch) so that its level-0 kernels store their tiles
    straight into the owner's memory.  Returns the owner list.  Collective: every rank must call it."""
    rank, world = dist.get_rank(), dist.get_world_size()
    owner = phenotype_owner(st.P, world)
    st.W_set_owned([1 if owner[p] == rank else 0 for p in range(st.P)])      # N x B x P/world of HBM per rank
    handles = [None] * world
    dist.all_gather_object(handles, st.W_export())
    for r in range(world):
        if r != rank:
            st.W_attach_peer(handles[r], [1 if owner[p] == r else 0 for p in range(st.P)])
    dist.barrier()
    return owner


def step1_distributed(st, n_blocks, feed_block, tau, chr_of_block, device, bt=None):
    """Step 1 with level-0 blocks sharded across ranks and level 1 sharded by phenotype.

    st          capi.Step1 created identically on every rank (total_blocks = n_blocks)
    feed_block  callable(block_id) -> None; runs rg_l0_block_bed for that block on this rank
    bt          None for quantitative traits, or (y_raw, offset) for the logistic level 1
    There is no data-path collective: each rank's level-0 kernels store the predictor tiles of a phenotype
    straight into the HBM of the rank that owns it (CUDA IPC mapping over NVLink/NVSwitch,
    rg_W_export / rg_W_attach_peer); torch.distributed carries the 64-byte handles, the barriers and the
    final gather of the P x R1 sums and N x 23 LOCO predictions.  Returns (cumsums, best_idx, loco) on every rank.
    """
    rank, world = dist.get_rank(), dist.get_world_size()
    attach_peers(st)
    first, n = partition_blocks(n_blocks, world)[rank]
    for b in range(first, first + n):
        feed_block(b)
    st.sync()
    code = first_error(int(st.status()), device if dist.get_backend() == "nccl" else None)
    if code:
        raise RuntimeError("level 0 failed on some rank (status %d)" % code)
    dist.barrier()                      # every rank's tiles have landed in their owners' HBM
    if bt is None:
        cs, best = st.l1_fit(tau)
    else:
        cs, best = st.l1_fit_bt(bt[0], bt[1], tau)
    loco = st.loco(chr_of_block)
    cs = _sum_to_all(cs, device)
    best = _sum_to_all(best.astype("int64"), device).astype("int32")
    loco = _sum_to_all(loco, device)
    dist.barrier()                      # peers may unmap / free W only after every owner is done
    return cs, best, loco


def step2_distributed(n_blocks, run_block):
    """Step 2 over more than one GPU: variants are independent (the reference splits Step-2 jobs by chromosome / variant
    range for the same reason), so the blocks are partitioned contiguously like level 0 and there is no data-path
    collective at all.  `run_block(b)` returns the finished result of block b (e.g. its `.regenie` rows); rank 0 gets the
    results of every block in file order, the other ranks get None.  Each rank keeps its own Step-2 handle / GPU."""
    mine = my_blocks(n_blocks)
    local = [(b, run_block(b)) for b in mine]
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [r for _, r in local]
    gathered = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if dist.get_rank() != 0:
        return None
    flat = sorted((b, r) for part in gathered for b, r in part)
    assert [b for b, _ in flat] == list(range(n_blocks))
    return [r for _, r in flat]
