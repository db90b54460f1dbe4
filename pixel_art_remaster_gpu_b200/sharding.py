"""How work is split across GPUs.  Pure host logic, no CUDA.

Frame batches: frames are independent units (the reference is stateless per launch_kernel call,
kernel.cu:313-523), so a stream is cut into contiguous per-rank shards and no collective touches the
data path.  Large single images: horizontal strips (row-major keeps a strip contiguous) plus an apron
of APRON_ROWS rows on each interior side — the exact dependency radius of the path is 37 source rows
(34 for the crossing walks, +2 for subdivision, +1 for the raster's 3x3 gather), NOT
the 1-pixel halo a plain stencil would need.
"""

APRON_ROWS = 40  # >= 37, rounded up so that strips stay 8-row aligned


def frame_shard(n_frames, rank, world):
    """[begin, end) of the contiguous shard of `n_frames` frames owned by `rank` (balanced to +-1)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    base, extra = divmod(n_frames, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def stream_seed(rank, frames_per_rank, base_seed):
    """First generator seed of a rank's frames when every rank synthesises its own shard (weak scaling)."""
    return base_seed + rank * frames_per_rank


def strip_rows(height, n_strips, apron=APRON_ROWS):
    """Per strip: (own_begin, own_end, load_begin, load_end) in image rows; own ranges tile [0, height),
    load ranges add the apron on interior sides (clipped at the image border)."""
    if n_strips < 1 or height < n_strips:
        raise ValueError("cannot cut %d rows into %d strips" % (height, n_strips))
    out = []
    for k in range(n_strips):
        b, e = frame_shard(height, k, n_strips)
        out.append((b, e, max(0, b - apron), min(height, e + apron)))
    return out


def max_over_ranks_ms(local_ms, dist=None, device=None):
    """Device time of a step = the slowest rank's (all_reduce MAX when a process group is up)."""
    if dist is None or not dist.is_available() or not dist.is_initialized():
        return float(local_ms)
    import torch
    t = torch.tensor([float(local_ms)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
