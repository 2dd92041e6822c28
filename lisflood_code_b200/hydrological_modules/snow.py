"""readmeteo / snow / frost / leafarea -- HydroModule mirrors of the feeder modules of a step (SURVEY.md §8 f3;
reference: hydrological_modules/readmeteo.py:44-81, snow.py:53-187, frost.py:44-78, leafarea.py:44-90).

On the device the scaling of the raw meteo maps, the three-zone snow model and the frost index are ONE kernel
(csrc/lf_model.cu::k_feeder, C ABI lf_model_feed) that reads the raw float32 / float64 maps of the step and leaves Rain,
SnowMelt, ETRef, EWRef, ESRef and isFrozenSoil where the soil stage reads them; only the raw maps cross PCIe.  The mirrors
keep the reference's call protocol: readmeteo.dynamic() takes the step's raw maps, snow.dynamic() runs the fused kernel,
frost.dynamic() checks the order, leafarea.dynamic() uploads the LAI maps when the 10-day interval changes.
`self.var` is a lisflood_code_b200.hotpath.HotPathModel."""
import numpy as np

from . import HydroModule

FEEDER_PARAMETERS = ("PrScaling", "CalEvaporation", "DeltaTSnow", "SnowSeason", "TempSnow", "SnowFactor", "SnowMeltCoef",
                     "TempMelt", "lat_rad", "Kfrost", "Afrost", "FrostIndexThreshold", "SnowWaterEquivalent", "kgb")
# snow.py:66-73
SNOW_DAY_DEGREES = 360 / 365.25
ICE_DAY_DEGREES = 2 * SNOW_DAY_DEGREES
ICEMELT_START_N, ICEMELT_END_N, ICEMELT_START_S, ICEMELT_END_S = 165, 257, 347, 74
# leafarea.py:50-51: first day of every interval over which the prescribed LAI is constant
LAI_INTERVAL_START = [1, 11, 21, 32, 42, 52, 60, 70, 80, 91, 101, 111, 121, 131, 141, 152, 162, 172, 182, 192, 202, 213, 223,
                      233, 244, 254, 264, 274, 284, 294, 305, 315, 325, 335, 345, 355, 370]


def season_coefficients(calendar_day):
    """The three scalars of snow.dynamic for a calendar day (snow.py:104-117): seasonal snow-melt coefficient and the
    summer ice-melt coefficient of the northern / southern hemisphere."""
    snowmelt_coeff = np.sin(np.radians((calendar_day - 81) * SNOW_DAY_DEGREES))
    is_summer_icemelt_N = (calendar_day > ICEMELT_START_N) & (calendar_day < ICEMELT_END_N)
    is_summer_icemelt_S = (calendar_day > ICEMELT_START_S) | (calendar_day < ICEMELT_END_S)
    _ice_melt_coeff = np.sin(np.radians((calendar_day - ICEMELT_START_N) * ICE_DAY_DEGREES))
    return float(snowmelt_coeff), float(_ice_melt_coeff if is_summer_icemelt_N else 0), \
        float(_ice_melt_coeff if is_summer_icemelt_S else 0)


def lai_interval(calendar_day):
    """Index of the 10-day LAI interval of a calendar day (the L1 lookup list of leafarea.py:63-69)."""
    j = 0
    for i in range(calendar_day + 1):
        if i >= LAI_INTERVAL_START[j + 1]:
            j += 1
    return j


class readmeteo(HydroModule):
    input_files_keys = {'all': ['PrecipitationMaps', 'TavgMaps', 'ET0Maps', 'E0Maps']}
    module_name = 'ReadMeteo'

    def __init__(self, readmeteo_variable):
        self.var = readmeteo_variable

    def dynamic(self, raw, calendar_day, asynchronous=False):
        """raw: {'Precipitation', 'Tavg', 'ET0', 'E0'} maps of the step as stored in the forcing files (float32 or
        float64, compressed order; NumPy arrays or torch tensors, host or CUDA).  The maps are only registered here; the
        scaling (readmeteo.py:66-69,78) is applied by the fused kernel that snow.dynamic() launches."""
        self.var.__dict__["_raw_meteo"] = (raw, int(calendar_day), bool(asynchronous))


class snow(HydroModule):
    input_files_keys = {'all': ['ElevationStD', 'TemperatureLapseRate', 'SnowSeasonAdj', 'TempSnow', 'SnowFactor', 'SnowMeltCoef',
                                'TempMelt', 'SnowCoverAInitValue', 'SnowCoverBInitValue', 'SnowCoverCInitValue']}
    module_name = 'Snow'

    def __init__(self, snow_variable):
        self.var = snow_variable

    def initial(self):
        """snow.py:53-93 on the host init object (lisflood_code_b200.Lisflood_initial.InitialVariables): parameters and the
        snow cover of the three elevation zones (-> HotPathModel.set_feeder through feeder_arguments below)."""
        v = self.var
        v.DeltaTSnow = 0.9674 * v.loadmap('ElevationStD') * v.loadmap('TemperatureLapseRate')   # :57
        v.SnowDayDegrees = SNOW_DAY_DEGREES
        v.IceDayDegrees = ICE_DAY_DEGREES
        v.SnowSeason = v.loadmap('SnowSeasonAdj') * 0.5                                         # :74
        v.TempSnow = v.loadmap('TempSnow')
        v.SnowFactor = v.loadmap('SnowFactor')
        v.SnowMeltCoef = v.loadmap('SnowMeltCoef')
        v.TempMelt = v.loadmap('TempMelt')
        init = [v.loadmap('SnowCover%sInitValue' % z) for z in "ABC"]
        v.SnowCoverS = init
        v.SnowCoverInit = (init[0] + init[1] + init[2]) / 3                                     # :88
        for k in ("SnowCover", "Snow", "Rain", "SnowMelt"):
            setattr(v, k, v.maskinfo.in_zero())

    def dynamic(self):
        pending = self.var.__dict__.pop("_raw_meteo", None)
        if pending is None:
            raise RuntimeError("snow.dynamic before readmeteo.dynamic of this step (Lisflood_dynamic.py:79-105)")
        raw, day, asynchronous = pending
        self.var.feed(raw, day, asynchronous=asynchronous)
        self.var.__dict__["_frost_pending"] = True


class frost(HydroModule):
    input_files_keys = {'all': ['Kfrost', 'Afrost', 'FrostIndexThreshold', 'SnowWaterEquivalent', 'FrostIndexInitValue']}
    module_name = 'Frost'

    def __init__(self, frost_variable):
        self.var = frost_variable

    def initial(self):
        """frost.py:44-58 on the host init object."""
        v = self.var
        v.Kfrost = v.loadmap('Kfrost')
        v.Afrost = v.loadmap('Afrost')
        v.FrostIndexThreshold = v.loadmap('FrostIndexThreshold')
        v.SnowWaterEquivalent = v.loadmap('SnowWaterEquivalent')
        v.FrostIndex = v.loadmap('FrostIndexInitValue')
        v.isFrozenSoil = v.FrostIndex > v.FrostIndexThreshold

    def dynamic(self):
        if not self.var.__dict__.pop("_frost_pending", False):
            raise RuntimeError("frost.dynamic before snow.dynamic of this step (Lisflood_dynamic.py:102-105)")


class leafarea(HydroModule):
    input_files_keys = {'all': ['kdf', 'LAIOtherMaps', 'LAIForestMaps', 'LAIIrrigationMaps']}
    module_name = 'LeafArea'

    def __init__(self, leafarea_variable):
        self.var = leafarea_variable
        self._interval = None

    def initial(self):
        """leafarea.py:44-72 on the host init object: extinction coefficient and the calendar-day -> interval lookup (the
        36 prescribed LAI maps per fraction are handed to dynamic() by the caller: map IO is out of scope)."""
        v = self.var
        v.kgb = 0.75 * v.loadmap('kdf')                                                        # :48
        v.L1 = [lai_interval(i) for i in range(367)]                                           # :63-69
        v.LAI = v.allocateVariableAllVegetation()                                              # :71

    def dynamic(self, calendar_day, lai_of_interval):
        """lai_of_interval(j) -> (3, N) LAI maps (Rainfed, Forest, Irrigated prescribed fractions) of interval j; called
        only when the interval changes (the maps then stay resident on the device, with LAITerm = exp(-kgb LAI))."""
        j = lai_interval(int(calendar_day))
        if j != self._interval:
            self.var.set_lai(lai_of_interval(j))
            self._interval = j


def feeder_arguments(var, lat_rad):
    """(parameters, initial state) for HotPathModel.set_feeder from a host init object on which miscInitial (PrScaling,
    CalEvaporation: miscInitial.py:142-143), snow.initial, frost.initial and leafarea.initial have run; lat_rad: latitude
    of the pixels [rad] (miscInitial.py:183-184 reads it from the NetCDF template)."""
    n = var.num_pixel
    full = lambda x: np.zeros(n) + x
    P = {k: getattr(var, k) for k in FEEDER_PARAMETERS if k != "lat_rad"}
    P["lat_rad"] = np.asarray(lat_rad, np.float64)
    state = {"SnowCoverS": np.stack([full(x) for x in var.SnowCoverS]), "FrostIndex": full(var.FrostIndex)}
    return P, state
