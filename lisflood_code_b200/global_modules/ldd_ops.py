"""NumPy restatement of the PCRaster ldd operators the hot path needs at INIT time (SURVEY.md §A.6;
call sites: reference hydrological_modules/routing.py:90-170, structures.py:51-59).  Host-side and
init-only; semantics per the PCRaster manual, on compressed (1-D) arrays."""
import numpy as np

_DR = {1: 1, 2: 1, 3: 1, 4: 0, 5: 0, 6: 0, 7: -1, 8: -1, 9: -1}
_DC = {1: -1, 2: 0, 3: 1, 4: -1, 5: 0, 6: 1, 7: -1, 8: 0, 9: 1}


def downstream_index(ldd_codes, land_mask):
    """int64[N]: compressed index of the downstream pixel, -1 for pits (code 5/0) and for pixels whose
    target falls off the map or off the mask (lddrepair semantics: they become pits)."""
    mask = np.asarray(land_mask, bool)
    rows, cols = mask.shape
    lp = -np.ones(mask.shape, np.int64)
    lp[mask] = np.arange(int(mask.sum()))
    rr, cc = np.nonzero(mask)
    code = np.asarray(ldd_codes).astype(np.int64)
    dr = np.array([0, 1, 1, 1, 0, 0, 0, -1, -1, -1])[code]
    dc = np.array([0, -1, 0, 1, -1, 0, 1, -1, 0, 1])[code]
    tr, tc = rr + dr, cc + dc
    ok = (tr >= 0) & (tr < rows) & (tc >= 0) & (tc < cols) & (code != 5) & (code != 0)
    out = -np.ones(code.size, np.int64)
    out[ok] = lp[tr[ok], tc[ok]]
    return out


def lddrepair_codes(ldd_codes, land_mask):
    """Codes with every link that leaves the mask turned into a pit (PCRaster lddrepair)."""
    ds = downstream_index(ldd_codes, land_mask)
    out = np.asarray(ldd_codes, np.float64).copy()
    out[ds < 0] = 5.0
    return out


def lddmask_codes(ldd_codes, keep, land_mask=None):
    """lddmask(ldd, keep): codes where `keep`, 0 (missing value -> isolated pit in kinematicWave,
    kinematic_wave_parallel.py:68) elsewhere.  With `land_mask`, kept cells that drain into a dropped cell, or out of
    the map / the mask, become pits (PCRaster keeps the result a sound ldd)."""
    codes = np.asarray(ldd_codes, np.float64)
    keep = np.asarray(keep, bool)
    out = np.where(keep, codes, 0.0)
    if land_mask is not None:
        ds = downstream_index(codes, land_mask)
        cut = keep & (ds >= 0) & ~keep[np.maximum(ds, 0)]
        cut |= keep & (ds < 0) & (codes != 5) & (codes != 0)
        out[cut] = 5.0
    return out


def topological_order(ds):
    """Pixels sorted so that every pixel precedes its downstream pixel (by hops to the outlet)."""
    n = ds.size
    hops = np.zeros(n, np.int64)
    cur = ds.copy()
    active = cur >= 0
    rounds = 0
    while active.any():
        rounds += 1
        if rounds > n:   # a pixel cannot be more than n-1 hops from its outlet: the map has a cycle (PCRaster: unsound ldd)
            from .errors import LisfloodError
            raise LisfloodError("the local drain direction map contains a cycle")
        hops[active] += 1
        cur[active] = ds[cur[active]]
        active = cur >= 0
    return np.argsort(-hops, kind="stable"), hops


def accuflux(ds, x):
    """Downstream-accumulated sum of x including the cell itself (PCRaster accuflux)."""
    order, hops = topological_order(ds)
    acc = np.asarray(x, np.float64).copy()
    hs = hops[order]
    # process level by level from the farthest pixels
    bounds = np.flatnonzero(np.diff(hs)) + 1
    start = 0
    for end in list(bounds) + [order.size]:
        p = order[start:end]
        p = p[ds[p] >= 0]
        np.add.at(acc, ds[p], acc[p])
        start = end
    return acc


def upstream_sum(ds, x):
    """PCRaster upstream(ldd, x): sum of x over the direct upstream neighbours."""
    out = np.zeros(ds.size, np.float64)
    has = ds >= 0
    np.add.at(out, ds[has], np.asarray(x, np.float64)[has])
    return out


def catchment_of_pits(ds):
    """catchment(ldd, nominal(uniqueid(pit(ldd)))) on compressed arrays (routing.py:162-164): every pixel gets the id
    (1, 2, ... in row-major order of the pits) of the pit it drains to."""
    ds = np.asarray(ds, np.int64)
    n = ds.size
    root = np.where(ds >= 0, ds, np.arange(n))
    while True:  # pointer jumping
        nxt = root[root]
        if np.array_equal(nxt, root):
            break
        root = nxt
    ids = np.zeros(n, np.int64)
    pits = np.flatnonzero(ds < 0)
    ids[pits] = np.arange(1, pits.size + 1)
    return ids[root]
