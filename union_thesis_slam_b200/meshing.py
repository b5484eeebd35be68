"""get_mesh / get_point_cloud support (reference grid_fusion.py:322-360, hash_fusion.py:465-507).

The reference delegates to ``skimage.measure.marching_cubes_lewiner`` (removed from scikit-image
>= 0.19 and absent from this image); no reference test or fixture pins mesh output.  Surface
extraction is outside the integration hot path.  When scikit-image is
importable its marching cubes is used on the downloaded volume; otherwise these calls raise.
"""
import numpy as np


def _marching_cubes(tsdf_vol):
    try:
        from skimage import measure
    except ImportError as exc:
        raise RuntimeError("get_mesh/get_point_cloud need scikit-image's marching cubes, which is not "
                           "installed; the TSDF volume itself is available from get_volume()") from exc
    fn = getattr(measure, "marching_cubes_lewiner", None) or measure.marching_cubes
    return fn(tsdf_vol, level=0)


def _vertex_colors(color_vol, verts, color_const):
    ind = np.round(verts).astype(int)
    rgb = color_vol[ind[:, 0], ind[:, 1], ind[:, 2]]
    b = np.floor(rgb / color_const)
    g = np.floor((rgb - b * color_const) / 256)
    r = rgb - b * color_const - g * 256
    return np.floor(np.asarray([r, g, b])).T.astype(np.uint8)


def mesh(tsdf_vol, color_vol, voxel_size, vol_origin, color_const=65536):
    verts, faces, norms, _ = _marching_cubes(tsdf_vol)
    colors = _vertex_colors(color_vol, verts, color_const)
    return verts * voxel_size + vol_origin, faces, norms, colors


def point_cloud(tsdf_vol, color_vol, voxel_size, vol_origin, color_const=65536):
    verts = _marching_cubes(tsdf_vol)[0]
    colors = _vertex_colors(color_vol, verts, color_const)
    return np.hstack([verts * voxel_size + vol_origin, colors])
