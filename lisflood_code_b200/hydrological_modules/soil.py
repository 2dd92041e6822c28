"""soil -- HydroModule mirror (reference: src/lisflood/hydrological_modules/soil.py:471-514, dynamic_perpixel)
plus the init-time parameter derivation of soil.initial (soil.py:109-228, 353-376) as a host-side helper."""
from collections import OrderedDict

import numpy as np

from . import HydroModule


def mualem(w_res, w_sat, genu_alpha, genu_n, genu_m, head_cm):
    """pressure2SoilMoistureFun (soil.py:30-35): storage [mm] at pressure head [cm]."""
    return w_res + (w_sat - w_res) / ((1 + (genu_alpha * head_cm) ** genu_n) ** genu_m)


def derive_layer_parameters(depth, theta_s, theta_res, lam, genu_alpha):
    """van Genuchten / Mualem storages of one layer (soil.py:160-228): arrays of any common shape."""
    gn = 1 + lam
    gm = lam / gn
    ws, wres = theta_s * depth, theta_res * depth
    return {"GenuM": gm, "GenuInvM": 1 / gm, "WS": ws, "WRes": wres, "WFC": mualem(wres, ws, genu_alpha, gn, gm, 100),
            "WPF3": mualem(wres, ws, genu_alpha, gn, gm, 1000), "WWP": mualem(wres, ws, genu_alpha, gn, gm, 15000),
            "PoreSpaceNotZero": np.logical_and(depth != 0, ws != 0)}


# inputs per soil layer: (attribute suffix, depth maps, binding-name index, has a separate forest map)
_LAYERS = (("1a", ("SoilDepth1", "SoilDepth1Forest"), "1", True), ("1b", ("SoilDepth2", "SoilDepth2Forest"), "2", True),
           ("2", ("SoilDepth3", "SoilDepth3Forest"), "3", False))


class soil(HydroModule):
    input_files_keys = {'all': ['SoilDepth1', 'SoilDepth1Forest', 'SoilDepth2', 'SoilDepth2Forest', 'SoilDepth3',
                                'SoilDepth3Forest', 'CourantCrit', 'LeafDrainageTimeConstant', 'AvWaterRateThreshold',
                                'MapCropCoef', 'MapForestCropCoef', 'MapIrrigationCropCoef', 'MapCropGroupNumber',
                                'MapForestCropGroupNumber', 'MapIrrigationCropGroupNumber', 'MapN', 'MapForestN',
                                'MapKSat1', 'MapKSat1Forest', 'MapKSat2', 'MapKSat2Forest', 'MapKSat3', 'MapLambda1',
                                'MapLambda1Forest', 'MapLambda2', 'MapLambda2Forest', 'MapLambda3', 'MapGenuAlpha1',
                                'MapGenuAlpha1Forest', 'MapGenuAlpha2', 'MapGenuAlpha2Forest', 'MapGenuAlpha3',
                                'MapThetaSat1', 'MapThetaSat1Forest', 'MapThetaSat2', 'MapThetaSat2Forest',
                                'MapThetaSat3', 'MapThetaRes1', 'MapThetaRes1Forest', 'MapThetaRes2',
                                'MapThetaRes2Forest', 'MapThetaRes3', 'ThetaInit1Value', 'ThetaForestInit1Value',
                                'ThetaIrrigationInit1Value', 'ThetaInit2Value', 'ThetaForestInit2Value',
                                'ThetaIrrigationInit2Value', 'ThetaInit3Value', 'ThetaForestInit3Value',
                                'ThetaIrrigationInit3Value', 'b_Xinanjiang', 'PowerPrefFlow', 'DSLRInitValue',
                                'DSLRForestInitValue', 'DSLRIrrigationInitValue', 'CumIntInitValue',
                                'CumIntForestInitValue', 'CumIntIrrigationInitValue', 'CumIntSealedInitValue',
                                'SMaxSealed'],
                        'drainedIrrigation': ['DrainedFraction'],
                        'simulatePF': ['HeadMax']}
    module_name = 'Soil'

    def __init__(self, soil_variable):
        self.var = soil_variable

    def initial(self):
        """Parameter derivation of the soil module (reference: hydrological_modules/soil.py:76-470, prescribed
        vegetation, no EPIC).  `self.var` is an InitialVariables (Lisflood_initial.py); on a device-resident
        HotPathModel the parameters are already in place and there is nothing to do."""
        v = self.var
        if not hasattr(v, "defsoil"):
            return
        load = v.loadmap
        zeros = v.maskinfo.in_zero
        rainfed, forest, irrigated = (v.vegetation.index(x) for x in v.PRESCRIBED_VEGETATION)
        # land-use fractions (:89-102): rice is handled as part of the rainfed fraction
        v.SoilFraction.values[rainfed] += v.RiceFraction
        for nm in ("Water", "Other", "Irrigation", "Forest", "DirectRunoff"):
            setattr(v, nm + "FractionBase", getattr(v, nm + "Fraction").copy())
        v.PermeableFraction = 1 - v.DirectRunoffFraction - v.WaterFraction
        # miscellaneous parameters (:115-128)
        v.CourantCrit = load('CourantCrit')
        v.LeafDrainageK = np.minimum(v.DtDay * (1 / load('LeafDrainageTimeConstant')), 1)
        v.AvWaterThreshold = load('AvWaterRateThreshold') * v.DtDay
        # land-use maps (:139-147)
        v.CropCoef = v.defsoil('MapCropCoef', 'MapForestCropCoef', 'MapIrrigationCropCoef')
        v.CropGroupNumber = v.defsoil('MapCropGroupNumber', 'MapForestCropGroupNumber', 'MapIrrigationCropGroupNumber')
        v.NManning = v.defsoil('MapN', 'MapForestN', 0.02, OrderedDict([v.dim_runoff, v.dim_pixel]))
        # van Genuchten / Mualem storages per layer (:109-113, 158-228)
        for lay, depth_maps, idx, has_forest in _LAYERS:
            def both(stem):
                return v.defsoil(stem + idx, stem + idx + 'Forest') if has_forest else v.defsoil(stem + idx)
            depth = v.defsoil(*depth_maps)
            ksat, lam, galpha = both('MapKSat'), both('MapLambda'), both('MapGenuAlpha')
            theta_s, theta_r = both('MapThetaSat'), both('MapThetaRes')
            genu_n = 1 + lam
            genu_m = lam / genu_n
            ws, wres = theta_s * depth, theta_r * depth
            setattr(v, "SoilDepth" + lay, depth)
            setattr(v, "KSat" + lay, ksat)
            setattr(v, "GenuM" + lay, genu_m)
            setattr(v, "GenuInvM" + lay, 1 / genu_m)
            setattr(v, "GenuInvN" + lay, 1 / genu_n)
            setattr(v, "GenuInvAlpha" + lay, 1 / galpha)
            setattr(v, "WS" + lay, ws)
            setattr(v, "WRes" + lay, wres)
            setattr(v, {"1a": "WS1WResa", "1b": "WS1WResb", "2": "WS2WRes"}[lay], ws - wres)
            setattr(v, "WFC" + lay, mualem(wres, ws, galpha, genu_n, genu_m, 100))       # pF 2
            if lay != "2":
                setattr(v, "WPF3" + lay[1], mualem(wres, ws, galpha, genu_n, genu_m, 1000))  # pF 3
            setattr(v, "WWP" + lay, mualem(wres, ws, galpha, genu_n, genu_m, 15000))     # pF 4.2
            setattr(v, "PoreSpaceNotZero" + lay, np.logical_and(depth != 0, ws != 0))
        v.SoilDepthTotal = v.SoilDepth1a + v.SoilDepth1b + v.SoilDepth2
        for nm in ("WS", "WRes", "WFC", "WWP"):
            setattr(v, nm + "1", getattr(v, nm + "1a") + getattr(v, nm + "1b"))
        v.WPF3 = v.WPF3a + v.WPF3b
        # initial soil moisture (:233-277): field capacity where the initial value is -9999, 0 without pore space
        names = {"1a": ('ThetaInit1Value', 'ThetaForestInit1Value', 'ThetaIrrigationInit1Value'),
                 "1b": ('ThetaInit2Value', 'ThetaForestInit2Value', 'ThetaIrrigationInit2Value'),
                 "2": ('ThetaInit3Value', 'ThetaForestInit3Value', 'ThetaIrrigationInit3Value')}
        for lay in ("1a", "1b", "2"):
            theta0 = v.allocateVariableAllVegetation()
            for iveg, nm in zip((rainfed, forest, irrigated), names[lay]):
                theta0.values[iveg] = load(nm)
            w = v.allocateVariableAllVegetation()
            depth, wfc, pore = (getattr(v, a + lay) for a in ("SoilDepth", "WFC", "PoreSpaceNotZero"))
            for veg, luse in v.VEGETATION_LANDUSE.items():
                iveg, iluse = v.vegetation.index(veg), v.SOIL_USES.index(luse)
                ini = np.where(theta0[iveg] == -9999, wfc[iluse], theta0[iveg] * depth[iluse])
                w[iveg] = np.where(pore[iluse], ini, 0)
            setattr(v, "W" + lay, w)
        v.W1 = v.W1a + v.W1b
        for nm in ("Sat1a", "Sat1b", "Sat1", "Sat2"):
            setattr(v, nm, v.allocateVariableAllVegetation())
        # Xinanjiang infiltration and preferential flow (:353-376): scalars are spread over the mask
        for attr, binding in (("b_Xinanjiang", 'b_Xinanjiang'), ("PowerPrefFlow", 'PowerPrefFlow')):
            x = load(binding)
            setattr(v, attr, zeros() + x if isinstance(x, float) else x)
        v.PowerInfPot = (v.b_Xinanjiang + 1) / v.b_Xinanjiang
        v.StoreMaxPervious = v.WS1 / (v.b_Xinanjiang + 1)
        # days since last rain, interception stores (:388-410)
        v.DSLR = v.allocateVariableAllVegetation()
        v.CumInterception = v.allocateVariableAllVegetation()
        for iveg, dslr, cum in ((rainfed, 'DSLRInitValue', 'CumIntInitValue'),
                                (forest, 'DSLRForestInitValue', 'CumIntForestInitValue'),
                                (irrigated, 'DSLRIrrigationInitValue', 'CumIntIrrigationInitValue')):
            v.DSLR[iveg] = load(dslr)
            v.DSLR[iveg] = np.where(v.DSLR[iveg] < 1, 1, v.DSLR[iveg])
            v.CumInterception[iveg] = load(cum)
        for nm in ("TotalPrecipitation", "TaCUM", "TaWB", "TaInterceptionCUM", "TaInterceptionWB", "ESActCUM", "ESActWB"):
            setattr(v, nm, zeros())
        # sealed soil (:446-461)
        v.CumInterSealed = load('CumIntSealedInitValue')
        v.SMaxSealed = load('SMaxSealed')
        v.DrainedFraction = load('DrainedFraction') if v.option('drainedIrrigation') else 0.0
        if v.option('simulatePF'):       # :464-469; pF0..2 themselves are produced on request (HotPathModel.suction_pf)
            v.HeadMax = load('HeadMax')

    def dynamic_perpixel(self):
        self.var._soil_stage_call("dynamic_perpixel")
