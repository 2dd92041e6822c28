"""surface_routing -- HydroModule mirror (reference: src/lisflood/hydrological_modules/surface_routing.py:115-212).
The three overland routers (Direct / Other / Forest on LddToChan) are solved in one level sweep on the device
(lf_model_surface_routing); the runoff components they consume were produced by the soil stage."""
from . import HydroModule


class surface_routing(HydroModule):
    input_files_keys = {'all': ['OFOtherInitValue', 'OFForestInitValue', 'OFDirectInitValue', 'Grad', 'GradMin',
                                'OFDepRef']}
    module_name = 'SurfaceRouting'

    def __init__(self, surface_routing_variable):
        self.var = surface_routing_variable

    def initial(self):
        pass

    def initialSecond(self):
        pass   # the routers are built by lf_model_create from LddToChan

    def dynamic(self):
        self.var._require_soil_stage_done()
        self.var.surface_routing()
