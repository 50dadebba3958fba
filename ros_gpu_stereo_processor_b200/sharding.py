"""Frame sharding across GPUs: independent stereo frames, frame i -> rank i mod N, no data-path
collective.  `torch.distributed` is only used for the start barrier, the MAX-reduction of the device-timed span and the
optional gather of per-frame digests; the backend is nccl on GPUs and gloo in the CPU tests."""


def frames_of_rank(n_frames, rank, world):
    """Indices of the frames rank `rank` processes (round robin, like one frame queue per GPU)."""
    return list(range(rank, n_frames, world))


def merge_in_frame_order(per_rank_results, world):
    """Re-orders results delivered per rank (lists in processing order) by global frame index."""
    n = sum(len(r) for r in per_rank_results)
    out = [None] * n
    for rank, res in enumerate(per_rank_results):
        for j, item in enumerate(res):
            out[rank + j * world] = item
    return out


def aggregate_throughput(frames_per_rank, ms, dist=None, device=None):
    """Whole-job frames/s: all ranks' frames over the slowest rank's device-timed span (MAX over ranks)."""
    total = frames_per_rank
    if dist is not None and dist.is_initialized():
        import torch
        t = torch.tensor([float(ms), float(frames_per_rank)], dtype=torch.float64, device=device or "cpu")
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, total = float(tmax[0]), float(tsum[1])
    return total / (ms * 1e-3), ms
