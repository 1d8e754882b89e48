"""Multi-GPU plumbing. The path shards by independent baseband stream: stream i -> rank i mod world, one process
per GPU, NO collective on the data path. torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for the
step barrier, the max-over-ranks timing and the gather of per-stream counters."""
import torch
import torch.distributed as dist


def streams_of_rank(n_streams, rank, world):
    """Round-robin ownership of independent streams."""
    return [s for s in range(n_streams) if s % world == rank]


def stream_seed(config_index, stream):
    """Seeds of the synthetic streams: 0xB200_0000 + config# * 16 + stream#."""
    return 0xB2000000 + config_index * 16 + stream


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum over ranks of a list of floats (step times)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def sum_over_ranks(values, device="cpu"):
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t]


def all_true(flag, device="cpu"):
    t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=device)
    if world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(int(t[0]))


def gather_counters(counters, device="cpu"):
    """Every rank's integer counters (samples, CADUs, RS stats ...) on every rank: list (per rank) of lists."""
    t = torch.tensor(list(counters), dtype=torch.int64, device=device)
    if world() == 1:
        return [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(world())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


def aggregate_throughput(samples_per_rank_step, steps, step_seconds_local, device="cpu"):
    """Whole-job samples/s: all ranks' samples over the slowest rank's time."""
    total = sum_over_ranks([samples_per_rank_step * steps], device)[0]
    slowest = max_over_ranks([step_seconds_local], device)[0]
    return total / slowest
