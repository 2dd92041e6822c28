"""lakes -- HydroModule mirror, initialisation part (reference: src/lisflood/hydrological_modules/lakes.py:52-196).

Lakes act inside the routing sub-step loop (`dynamic_inloop`, lakes.py:199-297: Modified Puls): on the device the rule
is evaluated by the channel wavefront itself (csrc/lf_model.cu::lake_substep), GPU-tested against goldens made by the
reference's own classes (tests/test_gpu_structures.py).  This file derives the parameters on the host (pinned bit for
bit to the reference's own `initial()`).  Lookup tables are two-column arrays: site id, value."""
import warnings

import numpy as np

from . import HydroModule
from .reservoir import lookupscalar
from ..global_modules.errors import LisfloodWarning


class lakes(HydroModule):
    input_files_keys = {'simulateLakes': ['LakeSites', 'TabLakeArea', 'TabLakeA', 'LakeMultiplier', 'LakeInitialLevelValue',
                                          'TabLakeAvNetInflowEstimate', 'PrevDischarge', 'LakePrevInflowValue',
                                          'LakePrevOutflowValue']}
    module_name = 'Lakes'

    def __init__(self, lakes_variable):
        self.var = lakes_variable

    def initial(self):
        v = self.var
        if not hasattr(v, "defsoil") or not v.option('simulateLakes') or v.option('InitLisflood'):
            return
        from ..global_modules import ldd_ops
        from ..global_modules.add1 import makenumpy
        load, zeros = v.loadmap, v.maskinfo.in_zero
        sites = np.array(load('LakeSites'), np.float64)
        sites[sites < 1] = 0
        sites[v.IsChannel == 0] = 0
        on = sites > 0
        v.LakeSitesCC = np.compress(on, sites)
        v.LakeIndex = np.nonzero(sites)[0]
        if v.LakeSitesCC.size == 0:
            warnings.warn(LisfloodWarning('There are no lakes. Lakes simulation won\'t run'))
            v.options['simulateLakes'] = False
            return
        v.IsStructureKinematic = np.where(on, True, v.IsStructureKinematic)
        dsk = ldd_ops.downstream_index(v.LddKinematic, v.maskinfo.land_mask)
        v.IsUpsOfStructureLake = (dsk >= 0) & on[np.maximum(dsk, 0)]      # pixels just upstream of lakes (:89)
        inflow_now = np.bincount(v.downstruct, weights=v.ChanQ)[v.LakeIndex]
        v.LakeAreaCC = np.compress(on, lookupscalar(v.loadtable('TabLakeArea'), sites))
        v.LakeSitesC2 = sites
        v.LakeACC = np.compress(on, lookupscalar(v.loadtable('TabLakeA'), sites) * load('LakeMultiplier'))
        level0 = load('LakeInitialLevelValue')
        cold = np.max(level0) == -9999
        if cold:   # S = LakeArea * sqrt(Q / a) from the estimated average net inflow (:109-117)
            v.LakeAvNetCC = np.compress(on, lookupscalar(v.loadtable('TabLakeAvNetInflowEstimate'), sites))
            storage = v.LakeAreaCC * np.sqrt(v.LakeAvNetCC / v.LakeACC)
            v.LakeLevelCC = storage / v.LakeAreaCC
            v.LakeInflowOldCC = inflow_now
        else:
            v.LakeLevelCC = np.compress(on, makenumpy(level0, v.maskinfo))
            storage = v.LakeAreaCC * v.LakeLevelCC
            v.LakeAvNetCC = np.compress(on, makenumpy(load('PrevDischarge'), v.maskinfo))
            v.LakeInflowOldCC = np.compress(on, makenumpy(load('LakePrevInflowValue'), v.maskinfo))
        v.LakeFactor = v.LakeAreaCC / (v.DtRouting * np.sqrt(v.LakeACC))
        v.LakeFactorSqr = np.square(v.LakeFactor)
        indicator = storage / v.DtRouting + v.LakeAvNetCC / 2
        prev_out = load('LakePrevOutflowValue')
        if np.max(prev_out) == -9999:
            v.LakeOutflowCC = np.square(-v.LakeFactor + np.sqrt(v.LakeFactorSqr + 2 * indicator))
        else:
            v.LakeOutflowCC = np.compress(on, makenumpy(prev_out, v.maskinfo))
        v.LakeStorageM3CC = storage.copy()
        v.LakeStorageM3BalanceCC = storage.copy()
        for full, cc in (("LakeStorageIniM3", storage), ("LakeLevel", v.LakeLevelCC), ("LakeInflowOld", v.LakeInflowOldCC),
                         ("LakeOutflow", v.LakeOutflowCC)):
            a = zeros()
            np.put(a, v.LakeIndex, cc)
            setattr(v, full, a)
        v.LakeStorageM3 = v.LakeStorageIniM3.copy()
        v.EWLakeCUMM3 = zeros()
        v.EWLakeWBM3 = zeros()

    def dynamic_inloop(self, NoRoutingExecuted):
        """The sub-step rule (lakes.py:199-297) runs inside the device channel wavefront (csrc/lf_model.cu: the structure pixel's own
        work item of every sub-step, lf_model_set_structures); the reference's call from routing.dynamic is kept as a
        protocol check only.  State and outputs are read back through the model object (ReservoirStorageM3, LakeLevel ...)."""
        if not 0 <= NoRoutingExecuted < self.var.NoRoutSteps:
            raise RuntimeError("lakes.dynamic_inloop: sub-step %d out of range" % NoRoutingExecuted)
