"""opensealed -- HydroModule mirror (reference: src/lisflood/hydrological_modules/opensealed.py:41-71);
part of the fused soil stage (see soilloop.py)."""
from . import HydroModule


class opensealed(HydroModule):
    input_files_keys = {'all': []}
    module_name = 'OpenSealed'

    def __init__(self, opensealed_variable):
        self.var = opensealed_variable

    def dynamic(self):
        self.var._soil_stage_call("opensealed")
