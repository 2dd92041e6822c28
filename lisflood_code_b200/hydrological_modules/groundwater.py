"""groundwater -- HydroModule mirror (reference: src/lisflood/hydrological_modules/groundwater.py:134-180);
part of the fused soil stage (see soilloop.py)."""
import numpy as np

from . import HydroModule


class groundwater(HydroModule):
    input_files_keys = {'all': ['UpperZoneTimeConstant', 'LowerZoneTimeConstant', 'LZInitValue', 'LZThreshold',
                                'UZInitValue', 'UZForestInitValue', 'UZIrrigationInitValue']}
    module_name = 'GroundWater'

    def __init__(self, groundwater_variable):
        self.var = groundwater_variable

    def initial(self):
        """Zone constants and initial storages (reference: hydrological_modules/groundwater.py:44-132).  `self.var` is an
        InitialVariables; nothing to do on a device-resident HotPathModel."""
        v = self.var
        if not hasattr(v, "defsoil"):
            return
        from ..global_modules.add1 import makenumpy
        load, zeros = v.loadmap, v.maskinfo.in_zero
        tc = []
        for name in ('UpperZoneTimeConstant', 'LowerZoneTimeConstant'):
            x = load(name)
            tc.append(zeros() + x if isinstance(x, float) else x)
        UpperZoneTimeConstant, LowerZoneTimeConstant = tc
        v.UpperZoneK = np.minimum(v.DtDay * (1 / UpperZoneTimeConstant), 1)
        v.LowerZoneK = np.minimum(v.DtDay * (1 / LowerZoneTimeConstant), 1)
        if v.option('InitLisflood'):
            guess = v.GwPerc - v.GwLoss
        else:
            guess = np.minimum(load('LZAvInflowMap'), v.GwPerc - v.GwLoss)   # from the pre-run; cannot exceed GwPerc
        guess = makenumpy(guess, v.maskinfo)
        LZSteady = guess * LowerZoneTimeConstant
        LZInitValue = load('LZInitValue')
        v.LZ = np.where(LZInitValue == -9999, LZSteady, LZInitValue)
        v.LZThreshold = load('LZThreshold')
        v.UZ = v.allocateVariableAllVegetation()
        for veg, name in zip(v.PRESCRIBED_VEGETATION, ('UZInitValue', 'UZForestInitValue', 'UZIrrigationInitValue')):
            v.UZ.values[v.vegetation.index(veg)][:] = load(name)
        for nm in ("GwLossCUM", "LZInflowCUM", "GwLossLZ", "LZOutflow", "LZAvInflow", "GwPercUZLZPixel", "GwLossPixel"):
            setattr(v, nm, zeros())
        v.GwPercUZLZ = v.allocateVariableAllVegetation()
        v.UZOutflow = v.allocateVariableAllVegetation()

    def dynamic(self):
        self.var._soil_stage_call("groundwater")
