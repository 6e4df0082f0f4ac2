"""Pair sharding for multi-GPU inference: image pairs are independent units (evaluate.py loops pairs
at batch 1), so rank r of N simply owns pairs r, r+N, r+2N, ... -- no data-path collective.  The
only communication is the final reduction of scalar metrics (sum of EPE, pair count, max time)."""
import torch
import torch.distributed as dist


def pairs_for_rank(num_pairs, rank, world_size):
    """Indices of the pairs rank `rank` processes (round-robin keeps ranks within one pair of each other)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    return list(range(rank, num_pairs, world_size))


def reduce_metrics(epe_sum, count, elapsed_s, device="cpu"):
    """All ranks -> (mean EPE over all pairs, total pairs, max elapsed).  Works on gloo and nccl."""
    t = torch.tensor([float(epe_sum), float(count)], dtype=torch.float64, device=device)
    m = torch.tensor([float(elapsed_s)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
    total = t[1].item()
    return (t[0].item() / total if total else float("nan")), int(total), m.item()
