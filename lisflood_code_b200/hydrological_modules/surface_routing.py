"""surface_routing -- HydroModule mirror (reference: src/lisflood/hydrological_modules/surface_routing.py:115-212).
The three overland routers (Direct / Other / Forest on LddToChan) are solved in one level sweep on the device
(lf_model_surface_routing); the runoff components they consume were produced by the soil stage."""
import numpy as np

from . import HydroModule


class surface_routing(HydroModule):
    input_files_keys = {'all': ['OFOtherInitValue', 'OFForestInitValue', 'OFDirectInitValue', 'Grad', 'GradMin',
                                'OFDepRef']}
    module_name = 'SurfaceRouting'

    def __init__(self, surface_routing_variable):
        self.var = surface_routing_variable

    def initial(self):
        """Overland-flow alpha of the three runoff classes and the initial overland storage / discharge (reference:
        hydrological_modules/surface_routing.py:43-95).  `self.var` is an InitialVariables; nothing to do on a
        device-resident HotPathModel."""
        v = self.var
        if not hasattr(v, "defsoil"):
            return
        from ..global_modules.add1 import makenumpy
        load = v.loadmap
        v.WaterDepth = v.maskinfo.in_zero()
        v.OFM3Other = makenumpy(load('OFOtherInitValue'), v.maskinfo)
        v.OFM3Forest = makenumpy(load('OFForestInitValue'), v.maskinfo)
        v.OFM3Direct = makenumpy(load('OFDirectInitValue'), v.maskinfo)
        Grad = np.maximum(load('Grad'), load('GradMin'))
        v.NoSubStepsOF = 1
        OFWettedPerimeter = v.PixelLength + 2 * v.MMtoM * load('OFDepRef')
        v.OFAlpha = (((v.NManning / np.sqrt(Grad)) ** v.Beta) * (OFWettedPerimeter ** v.AlpPow)).astype(float)
        v.InvOFAlpha = 1 / v.OFAlpha
        runoff = v.dim_runoff[1]
        for nm in ("Direct", "Other", "Forest"):
            m3 = getattr(v, "OFM3" + nm)
            setattr(v, "OFQ" + nm, ((m3 * v.InvPixelLength * v.InvOFAlpha.values[runoff.index(nm)]) ** (v.InvBeta)).astype(float))

    def initialSecond(self):
        pass   # the routers are built by lf_model_create from LddToChan

    def dynamic(self):
        self.var._require_soil_stage_done()
        self.var.surface_routing()
