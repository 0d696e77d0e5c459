"""Multi-GPU partitioning of the hot path: the path shards by texture / face / mip level
with NO data-path collective -- every unit is an independent launch on one GPU.  One process per GPU;
torch.distributed is used only for the barrier and the max-over-ranks of device time.

`partition_units` is the longest-processing-time-first assignment the reference would get from its
`block_index % (threads + 1)` striding if units had equal cost; here costs are block counts, which is
what the kernels' time is proportional to."""
import heapq


def unit_costs(level_dims, faces=1):
    """Cost (4x4 block count) of every (face, level) unit of one texture."""
    return [((w + 3) // 4) * ((h + 3) // 4) for _ in range(faces) for (w, h) in level_dims]


def partition_units(costs, world_size):
    """Returns world_size lists of unit indices; deterministic, balanced (LPT), every unit exactly once."""
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out = [[] for _ in range(world_size)]
    for i in sorted(range(len(costs)), key=lambda i: (-costs[i], i)):
        load, r = heapq.heappop(heap)
        out[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return [sorted(u) for u in out]


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0])


def allgather_inplace(buf_u8, bytes_per_rank, device=None):
    """In-place all-gather of a numpy uint8 buffer of world * bytes_per_rank bytes whose own slice is filled (the exchange
    crn_gpu_hc_compress asks for when one texture is sharded over the ranks).  gloo gathers the host tensor directly; NCCL
    needs device tensors, so the (small: 16 bytes per cluster) buffer is staged through `device`."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    full = torch.from_numpy(buf_u8).view(world, bytes_per_rank)
    if dist.get_backend() == "nccl":
        d = full.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))
        dist.all_gather_into_tensor(d.view(-1), d[rank].clone())
        full.copy_(d.cpu())
    else:
        out = [torch.empty(bytes_per_rank, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(out, full[rank].clone())
        for r in range(world):
            full[r].copy_(out[r])
