"""PCRaster-style map API kept at the hot-path boundary
(reference: src/lisflood/global_modules/add1.py:48-61,268-305 and settings.py:235-251).

Model state is held as 1-D float64 arrays of the active ("unmasked") pixels in row-major order of the
mask; 2-D (vegetation|landuse, pixel) arrays carry `.values` / `.dims` like the reference's NumpyModified.
"""
import numpy as np


class NumpyModified(np.ndarray):
    """ndarray with `.values` (itself) and `.dims` -- same surface as the reference class
    (add1.py:48-61), so code written against xarray-like attributes keeps working."""
    obj_dims = []

    def __new__(cls, input_array, dims):
        obj = np.asarray(input_array).view(cls)
        obj.obj_dims = list(dims)
        return obj

    def __array_finalize__(self, obj):
        if obj is not None:
            self.obj_dims = getattr(obj, "obj_dims", [])

    @property
    def values(self):
        return self

    @property
    def dims(self):
        return self.obj_dims


class MaskInfo(object):
    """Holds the model mask.  `mask` follows the reference convention: True = pixel EXCLUDED
    (settings.py:235-251); `land_mask` is its complement."""
    _instance = None

    def __init__(self, mask_excluded):
        self.mask = np.ascontiguousarray(mask_excluded, bool)
        self.shape = self.mask.shape
        self.land_mask = ~self.mask
        self.num_pixels = int(self.land_mask.sum())
        self.mapC = (self.num_pixels,)
        MaskInfo._instance = self

    @classmethod
    def instance(cls):
        if cls._instance is None:
            raise RuntimeError("MaskInfo not initialised")
        return cls._instance

    def in_zero(self):
        return np.zeros(self.num_pixels, np.float64)


def compressArray(map2d, maskinfo=None):
    """2-D map -> float64[N] of the active pixels (add1.py:268-282)."""
    mi = maskinfo or MaskInfo.instance()
    a = np.asarray(map2d)
    if a.shape != mi.shape:
        raise ValueError("map shape %s does not match the mask %s" % (a.shape, mi.shape))
    return np.ascontiguousarray(a[mi.land_mask])


def decompress(values, maskinfo=None, fill=-9999.0):
    """float64[N] -> 2-D map with `fill` outside the mask (add1.py:285-305)."""
    mi = maskinfo or MaskInfo.instance()
    out = np.full(mi.shape, fill, np.float64)
    out[mi.land_mask] = values
    return out


def makenumpy(v, maskinfo=None):
    """scalar or map -> float64[N] (add1.py makenumpy)."""
    mi = maskinfo or MaskInfo.instance()
    if np.ndim(v) == 0:
        return np.full(mi.num_pixels, float(v))
    return np.ascontiguousarray(v, np.float64)


def loadmap(name, binding=None, maskinfo=None):
    """Static input by binding name with the reference's return convention (add1.py:318-541): a Python float
    when the binding is a number, else a compressed float64[N] array -- several modules branch on
    `isinstance(x, float)` (soil.py:355,372; groundwater.py:54,60).  Map files here are NumPy `.npy` archives
    (2-D rasters are compressed with the mask, 1-D arrays are taken as already compressed); PCRaster / NetCDF
    readers are out of scope."""
    import os
    if binding is None:
        from .settings import LisSettings
        binding = LisSettings.instance().binding
    value = binding[name]
    try:
        return float(value)
    except (TypeError, ValueError):
        pass
    if not os.path.exists(value):
        raise FileNotFoundError("binding %s -> %s" % (name, value))
    a = np.load(value)
    if a.ndim == 2:
        return compressArray(a, maskinfo).astype(np.float64)
    return np.ascontiguousarray(a, np.float64)
