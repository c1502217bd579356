"""nind_denoise_b200 — B200-native (sm_100a) tiled inference of the NIND UtNet/UNet denoisers.

Drop-in for the hot path of esq4/nind-denoise (crop gather -> network forward -> trim/seam/stitch):
``UtNet`` / ``UNet`` are ``nn.Module``s with the reference's constructor, ``forward`` and
``state_dict`` layout; ``denoise_tiled`` is the reference's crop/stitch loop.  Everything numerical
runs in hand-written CUDA kernels behind the C ABI in include/nind_b200.h (libnind_b200.so).
"""
from .networks import UNet, UtNet, register  # noqa: F401
from .tiler import (CS_UNET, CS_UTNET, UCS_UNET, UCS_UTNET, assemble_bands, crop_table, denoise_tiled,  # noqa: F401
                    denoise_images_host, denoise_tiled_distributed, denoise_tiled_distributed_host, denoise_tiled_host, gather_crops, n_crops, shard_ranges,
                    stitch_crops, OneImageDS, SharedHostImage, PeerGather, plan_steps, tiled_step, add_rows, band_extents, denoise_whole_image, pad_whole_image, owned_rows, exchange_seams, owned_rows_up, seam_plan, exchange_seams_up, PeerSeams, bind_host_to_gpu)

__all__ = ["UtNet", "UNet", "register", "denoise_tiled", "denoise_tiled_host", "denoise_images_host", "denoise_tiled_distributed", "denoise_tiled_distributed_host",
           "crop_table", "n_crops", "shard_ranges", "assemble_bands", "gather_crops", "stitch_crops", "OneImageDS",
           "SharedHostImage", "PeerGather", "plan_steps", "tiled_step", "add_rows", "band_extents", "denoise_whole_image", "pad_whole_image", "owned_rows", "exchange_seams", "owned_rows_up", "seam_plan", "exchange_seams_up", "PeerSeams", "bind_host_to_gpu"]
