"""Per-step call order of the hot path (reference: src/lisflood/Lisflood_dynamic.py:114-229), driving the
HydroModule mirrors on a device-resident HotPathModel.  The feeder modules of the step (readmeteo, leafarea, snow, frost;
Lisflood_dynamic.py:79-105) run on the device when the step is given the RAW meteo maps (`raw=`), or are bypassed when
their products are handed in (`forcing=`: Rain, SnowMelt, ETRef, EWRef, ESRef, LAI, LAITerm, isFrozenSoil)."""
from .hydrological_modules.groundwater import groundwater
from .hydrological_modules.opensealed import opensealed
from .hydrological_modules.routing import routing
from .hydrological_modules.soil import soil
from .hydrological_modules.snow import frost, leafarea, readmeteo, snow
from .hydrological_modules.soilloop import soilloop
from .hydrological_modules.surface_routing import surface_routing


class LisfloodModel_dyn(object):
    def __init__(self, var):
        self.var = var
        self.soilloop_module = soilloop(var)
        self.soil_module = soil(var)
        self.opensealed_module = opensealed(var)
        self.groundwater_module = groundwater(var)
        self.surface_routing_module = surface_routing(var)
        self.routing_module = routing(var)
        self.readmeteo_module = readmeteo(var)
        self.leafarea_module = leafarea(var)
        self.snow_module = snow(var)
        self.frost_module = frost(var)
        self.NoRoutSteps = var.NoRoutSteps

    def dynamic(self, forcing=None, raw=None, calendar_day=None, lai_of_interval=None, asynchronous=False):
        if raw is not None:
            self.readmeteo_module.dynamic(raw, calendar_day, asynchronous)          # :79
            if lai_of_interval is not None:
                self.leafarea_module.dynamic(calendar_day, lai_of_interval)         # :95
            self.snow_module.dynamic()                                              # :102
            self.frost_module.dynamic()                                             # :105
        else:
            self.var.set_forcing(forcing)
        self.soilloop_module.dynamic_canopy()        # :114
        self.soilloop_module.dynamic_soil()          # :123
        self.opensealed_module.dynamic()             # :129
        self.soil_module.dynamic_perpixel()          # :147
        self.groundwater_module.dynamic()            # :149
        self.surface_routing_module.dynamic()        # :165
        for NoRoutingExecuted in range(self.NoRoutSteps):   # :179-180
            self.routing_module.dynamic(NoRoutingExecuted)
