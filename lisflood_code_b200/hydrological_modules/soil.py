"""soil -- HydroModule mirror (reference: src/lisflood/hydrological_modules/soil.py:471-514, dynamic_perpixel)
plus the init-time parameter derivation of soil.initial (soil.py:109-228, 353-376) as a host-side helper."""
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


class soil(HydroModule):
    input_files_keys = {'all': []}
    module_name = 'Soil'

    def __init__(self, soil_variable):
        self.var = soil_variable

    def initial(self):
        pass

    def dynamic_perpixel(self):
        self.var._soil_stage_call("dynamic_perpixel")
