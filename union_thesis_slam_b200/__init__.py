"""B200-native TSDF integration (dense brick grid + voxel-block hash map).

Drop-in for the mapping hot path of DiWu9/Union-Thesis-SLAM:

    from union_thesis_slam_b200 import grid_fusion, hash_fusion
    vol = grid_fusion.TSDFVolume(vol_bnds, voxel_size=0.02)
    vol.integrate(color_im, depth_im, cam_intr, cam_pose, obs_weight=1.)
    tsdf, color = vol.get_volume()

All voxel work runs in hand-written sm_100a CUDA behind the C ABI of include/tsdf_b200.h.
"""
from . import _capi  # noqa: F401
from . import grid_fusion, hash_fusion  # noqa: F401
from .grid_fusion import TSDFVolume  # noqa: F401
from .hash_fusion import HashTable  # noqa: F401
