"""soilloop -- HydroModule mirror (reference: src/lisflood/hydrological_modules/soilloop.py:437-704).

`self.var` is a lisflood_code_b200.hotpath.HotPathModel (the device-resident model object).  The reference runs
canopy, soil columns, open/sealed, per-pixel sums and groundwater as five module calls
(Lisflood_dynamic.py:114-149); on the device they are ONE fused stage (lf_model_soil), executed by the first of
the five calls of a step -- the others check the call order and return, because nothing of the hot path can
observe the intermediate states (the optional modules that could -- rice, water abstraction -- are out of scope).
"""
from . import HydroModule


class soilloop(HydroModule):
    input_files_keys = {'wateruse': []}
    module_name = 'SoilLoop'

    def __init__(self, soilloop_variable):
        self.var = soilloop_variable

    def initial(self):
        pass

    def dynamic_canopy(self):
        self.var._soil_stage_call("dynamic_canopy")

    def dynamic_soil(self):
        self.var._soil_stage_call("dynamic_soil")
