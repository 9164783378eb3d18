"""Frame-wise data parallelism: frames / streams are independent, so a batch is split
contiguously over ranks with no data-path collective; the only exchange is the gather of the small
fixed-size result records (torch.distributed: NCCL on GPUs, gloo in the CPU tests)."""


def shard_range(n_items, rank, world):
    """Contiguous split: rank g owns [g*n/N, (g+1)*n/N)."""
    return (rank * n_items) // world, ((rank + 1) * n_items) // world


def gather_records(local, world_counts, group=None):
    """all_gather of per-frame records (rows of a 2-D tensor) whose per-rank row counts may differ.

    local: (n_local, R) tensor; world_counts: list of n_local for every rank.  Returns (sum, R)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    mx = max(world_counts)
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([out[r][: world_counts[r]] for r in range(world)], dim=0)
