"""Hand partition of one render over `world` ranks through the plain C-ABI fields (include/ptc.h: split_mode, rank, world,
tile_size) - what a launcher without a communicator does; contexts that own several GPUs or a communicator fill these in
themselves (ptc_create with n_devices > 1, ptc_comm_init_rank)."""
from vviewer_b200 import capi

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
