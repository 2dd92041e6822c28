"""reservoir -- HydroModule mirror, initialisation part (reference: src/lisflood/hydrological_modules/reservoir.py:52-170).

Reservoirs act inside the routing sub-step loop (`dynamic_inloop`, reservoir.py:173-322): on the device the
four-regime outflow rule is evaluated by the channel wavefront itself (csrc/lf_model.cu::reservoir_substep), GPU-tested
against goldens made by the reference's own classes (tests/test_gpu_structures.py).  This file derives the parameters on
the host (pinned bit for bit to the reference's own `initial()`).
PCRaster lookup tables (`TabTotStorage` ...) are two-column arrays here: site id, value."""
import warnings

import numpy as np

from . import HydroModule
from ..global_modules.errors import LisfloodWarning


def lookupscalar(table, sites):
    """PCRaster lookupscalar(table, nominal map) on compressed arrays: value of the row whose key equals the site id,
    NaN where the id is 0 / missing or has no row."""
    table = np.asarray(table, np.float64).reshape(-1, 2)
    out = np.full(np.shape(sites), np.nan)
    ids = np.asarray(sites)
    for key, val in table:
        out[ids == key] = val
    return out


class reservoir(HydroModule):
    input_files_keys = {'simulateReservoirs': ['ReservoirSites', 'TabTotStorage', 'TabConservativeStorageLimit',
                                               'TabNormalStorageLimit', 'TabFloodStorageLimit', 'TabNonDamagingOutflowQ',
                                               'TabNormalOutflowQ', 'TabMinOutflowQ', 'adjust_Normal_Flood',
                                               'ReservoirRnormqMult', 'ReservoirInitialFillValue']}
    module_name = 'Reservoir'

    def __init__(self, reservoir_variable):
        self.var = reservoir_variable

    def initial(self):
        v = self.var
        if not hasattr(v, "defsoil") or not v.option('simulateReservoirs') or v.option('InitLisflood'):
            return
        from ..global_modules.add1 import makenumpy
        load, zeros = v.loadmap, v.maskinfo.in_zero
        sites = np.array(load('ReservoirSites'), np.float64)
        sites[sites < 1] = 0
        sites[v.IsChannel == 0] = 0          # reservoirs off the channel network are dropped (:70-72)
        v.ReservoirSitesC = sites
        v.ReservoirSitesCC = np.compress(sites > 0, sites)
        if v.ReservoirSitesCC.size == 0:
            warnings.warn(LisfloodWarning('There are no reservoirs. Reservoirs simulation won\'t run'))
            v.options['simulateReservoirs'] = False
            return
        on = sites > 0
        v.ReservoirIndex = np.nonzero(sites)[0]
        v.IsStructureKinematic = np.where(on, True, v.IsStructureKinematic)

        def table(name):
            return np.compress(on, lookupscalar(v.loadtable(name), sites))
        total = lookupscalar(v.loadtable('TabTotStorage'), sites)
        v.TotalReservoirStorageM3C = np.where(np.isnan(total), 0, total)
        v.TotalReservoirStorageM3CC = np.compress(on, v.TotalReservoirStorageM3C)
        v.ConservativeStorageLimitCC = table('TabConservativeStorageLimit')
        v.NormalStorageLimitCC = table('TabNormalStorageLimit')
        v.FloodStorageLimitCC = table('TabFloodStorageLimit')
        v.NonDamagingReservoirOutflowCC = table('TabNonDamagingOutflowQ')
        v.NormalReservoirOutflowCC = table('TabNormalOutflowQ')
        v.MinReservoirOutflowCC = table('TabMinOutflowQ')
        # calibration (:128-145)
        adjust = np.compress(on, makenumpy(load('adjust_Normal_Flood'), v.maskinfo))
        v.Normal_FloodStorageLimitCC = v.NormalStorageLimitCC + adjust * (v.FloodStorageLimitCC - v.NormalStorageLimitCC)
        mult = np.compress(on, makenumpy(load('ReservoirRnormqMult'), v.maskinfo))
        norm = v.NormalReservoirOutflowCC * mult
        norm = np.where(norm > v.MinReservoirOutflowCC, norm, v.MinReservoirOutflowCC + 0.01)
        v.NormalReservoirOutflowCC = np.where(norm < v.NonDamagingReservoirOutflowCC, norm, v.NonDamagingReservoirOutflowCC - 0.01)
        v.DeltaO = v.NormalReservoirOutflowCC - v.MinReservoirOutflowCC
        v.DeltaLN = v.NormalStorageLimitCC - 2 * v.ConservativeStorageLimitCC
        v.DeltaLF = v.FloodStorageLimitCC - v.NormalStorageLimitCC
        v.DeltaNFL = v.FloodStorageLimitCC - v.Normal_FloodStorageLimitCC
        # initial fill: -9999 = filled to the normal storage limit (:153-170)
        fill0 = load('ReservoirInitialFillValue')
        fill = v.NormalStorageLimitCC if np.max(fill0) == -9999 else np.compress(on, makenumpy(fill0, v.maskinfo))
        v.ReservoirFillCC = fill
        storage = fill * v.TotalReservoirStorageM3CC
        v.ReservoirStorageM3CC = storage.copy()
        v.ReservoirFill = zeros()
        v.ReservoirStorageIniM3 = zeros()
        np.put(v.ReservoirStorageIniM3, v.ReservoirIndex, storage)
        v.ReservoirStorageM3 = v.ReservoirStorageIniM3

    def dynamic_inloop(self, NoRoutingExecuted):
        """The sub-step rule (reservoir.py:173-322) runs inside the device channel wavefront (csrc/lf_model.cu: the structure pixel's own
        work item of every sub-step, lf_model_set_structures); the reference's call from routing.dynamic is kept as a
        protocol check only.  State and outputs are read back through the model object (ReservoirStorageM3, LakeLevel ...)."""
        if not 0 <= NoRoutingExecuted < self.var.NoRoutSteps:
            raise RuntimeError("reservoirs.dynamic_inloop: sub-step %d out of range" % NoRoutingExecuted)
