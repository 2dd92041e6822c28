"""groundwater -- HydroModule mirror (reference: src/lisflood/hydrological_modules/groundwater.py:134-180);
part of the fused soil stage (see soilloop.py)."""
from . import HydroModule


class groundwater(HydroModule):
    input_files_keys = {'all': ['UpperZoneTimeConstant', 'LowerZoneTimeConstant', 'LZInitValue', 'LZThreshold',
                                'UZInitValue', 'UZForestInitValue', 'UZIrrigationInitValue']}
    module_name = 'GroundWater'

    def __init__(self, groundwater_variable):
        self.var = groundwater_variable

    def initial(self):
        pass

    def dynamic(self):
        self.var._soil_stage_call("groundwater")
