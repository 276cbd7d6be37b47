"""Multi-GPU partitioning of one render: the scene is replicated, the image is cut into tiles or the
sample batches are dealt round-robin, and the per-rank accumulation buffers are SUMMED (tiles are disjoint, so one
reduce serves both modes).  One process per GPU; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) is plumbing."""
from . import capi

MODES = {"none": capi.PTC_SPLIT_NONE, "tile": capi.PTC_SPLIT_TILE, "sample": capi.PTC_SPLIT_SAMPLE}


def partition(rp, rank, world, mode, tile_size=32):
    """Fill the partition fields of a ptc_render_params for this rank (in place) and return it."""
    if world <= 1 or mode == "none":
        rp.split_mode, rp.rank, rp.world = capi.PTC_SPLIT_NONE, 0, 1
        return rp
    rp.split_mode = MODES[mode]
    rp.rank, rp.world, rp.tile_size = rank, world, tile_size
    return rp


def batches_of_rank(samples, batch_size, rank, world, mode):
    """How many batches this rank renders (sample split deals batch b to rank b % world)."""
    batches = samples // batch_size
    if mode != "sample" or world <= 1:
        return batches
    return len(range(rank, batches, world))


def reduce_to_root(dist, tensor, root=0):
    """Sum the accumulation buffers onto the root (NCCL reduce over NVLink on GPUs)."""
    dist.reduce(tensor, dst=root)
    return tensor
