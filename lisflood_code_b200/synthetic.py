"""Seeded synthetic catchments for the parity tests and bench.py.

Everything here is host-side NumPy input generation; it is not part of the timed path.
LDD codes are the PCRaster keypad codes the reference consumes
(reference: src/lisflood/hydrological_modules/kinematic_wave_parallel.py:47-51):

      7 8 9        row 0 = north, so 8 = (row-1, col), 2 = (row+1, col), 6 = (row, col+1) ...
      4 5 6        5 = pit
      1 2 3
"""
import numpy as np

# (drow, dcol) -> keypad code
_NEIGH = [(-1, -1, 7), (-1, 0, 8), (-1, 1, 9), (0, -1, 4), (0, 1, 6), (1, -1, 1), (1, 0, 2), (1, 1, 3)]


def random_ldd(rows, cols, seed=0, noise=3.0, tilt=1.0, mask_fraction=0.0):
    """Random D8 drainage forest: elevation = tilted plane + Gaussian noise, every cell drains to
    its steepest-descent neighbour, a cell with no strictly lower neighbour is a pit (code 5).

    noise/tilt ~ 3   -> shallow forest (tens of levels, many pits)
    noise/tilt ~ 0.3 -> deep trees (about `rows` levels)
    Returns (ldd_codes float64[rows, cols], land_mask bool[rows, cols]); cells outside the mask
    carry code 0.
    """
    rng = np.random.default_rng(seed)
    elev = rng.standard_normal((rows, cols)) * noise
    elev += tilt * np.arange(rows, dtype=np.float64)[::-1, None]          # drains to the south edge
    elev += 0.05 * tilt * np.abs(np.arange(cols, dtype=np.float64) - cols / 2)[None, :]  # towards a trunk valley
    if mask_fraction > 0:
        land = rng.random((rows, cols)) >= mask_fraction
    else:
        land = np.ones((rows, cols), bool)
    big = np.float64(1e300)
    e = np.where(land, elev, big)
    pad = np.full((rows + 2, cols + 2), big)
    pad[1:-1, 1:-1] = e
    best = np.zeros((rows, cols))
    code = np.full((rows, cols), 5.0)
    for dr, dc, k in _NEIGH:
        nb = pad[1 + dr:1 + dr + rows, 1 + dc:1 + dc + cols]
        dist = np.sqrt(float(dr * dr + dc * dc))
        drop = (e - nb) / dist
        drop[nb >= big] = -1.0
        better = drop > best
        best = np.where(better, drop, best)
        code = np.where(better, float(k), code)
    code[~land] = 0.0
    return code, land


def routing_fields(n, seed=0):
    """alpha, Q0, q for config C2: alpha~U(0.5,3), Q0~U(0.1,10), q~U(0,1e-4)."""
    rng = np.random.default_rng(seed + 7919)
    alpha = rng.uniform(0.5, 3.0, n)
    q0 = rng.uniform(0.1, 10.0, n)
    q = rng.uniform(0.0, 1e-4, n)
    return alpha, q0, q


def level_stats(order_start_stop):
    sizes = order_start_stop[:, 1] - order_start_stop[:, 0]
    return {"levels": int(sizes.size), "median_level": float(np.median(sizes)), "max_level": int(sizes.max()),
            "min_level": int(sizes.min())}
