"""CPU restatement (NumPy float64) of the feeder modules of a model step.  TEST INFRASTRUCTURE ONLY.

  readmeteo scaling  hydrological_modules/readmeteo.py:61-81
  snow               hydrological_modules/snow.py:95-187
  frost              hydrological_modules/frost.py:61-78
  LAITerm            hydrological_modules/leafarea.py:48,90
Pinned by tests/test_oracle_feeders_golden.py to goldens made by the reference's own snow / frost classes
(tests/golden/make_golden.py::feeders_case)."""
import numpy as np

ICE_START_N, ICE_END_N, ICE_START_S, ICE_END_S = 165, 257, 347, 74      # snow.py:70-71
SNOW_DAY_DEGREES = 360 / 365.25                                         # snow.py:66
ICE_DAY_DEGREES = 2 * SNOW_DAY_DEGREES                                  # snow.py:73


def season_coefficients(calendar_day):
    """(snowmelt_coeff, ice_melt_coeff_N, ice_melt_coeff_S): the three scalars of snow.py:104-118."""
    c = np.sin(np.radians((calendar_day - 81) * SNOW_DAY_DEGREES))
    ice = np.sin(np.radians((calendar_day - ICE_START_N) * ICE_DAY_DEGREES))
    in_n = (calendar_day > ICE_START_N) & (calendar_day < ICE_END_N)
    in_s = (calendar_day > ICE_START_S) | (calendar_day < ICE_END_S)
    return float(c), float(ice if in_n else 0), float(ice if in_s else 0)


class FeederOracle(object):
    def __init__(self, P, state, dt_sec):
        self.P = {k: np.asarray(v, np.float64) for k, v in P.items()}
        self.DtDay = dt_sec / 86400.0
        self.SnowCoverS = [np.array(state["SnowCoverS"][i], np.float64) for i in range(3)]
        self.FrostIndex = np.array(state["FrostIndex"], np.float64)
        self.TotalPrecipitation = np.zeros_like(self.FrostIndex)

    def step(self, raw, calendar_day):
        P, dt = self.P, self.DtDay
        o = {}
        prec = np.asarray(raw["Precipitation"], np.float64) * dt * P["PrScaling"]          # readmeteo.py:66
        tavg = np.asarray(raw["Tavg"], np.float64)
        o["Precipitation"] = prec
        o["ETRef"] = np.asarray(raw["ET0"], np.float64) * dt * P["CalEvaporation"]
        o["EWRef"] = np.asarray(raw["E0"], np.float64) * dt * P["CalEvaporation"]
        o["ESRef"] = (o["EWRef"] + o["ETRef"]) / 2
        c, ice_n, ice_s = season_coefficients(calendar_day)
        north = P["lat_rad"] > 0
        seas = P["SnowSeason"] * np.where(north, c, -c) + P["SnowMeltCoef"]                  # snow.py:107
        summer = np.where(north, ice_n, ice_s)                                               # :118
        snow, rain, melt, cover = (np.zeros_like(prec) for _ in range(4))
        for i in range(3):
            tz = tavg + P["DeltaTSnow"] * (i - 1)                                            # :151
            snow_s = np.where(tz < P["TempSnow"], P["SnowFactor"] * prec, 0.0)
            rain_s = np.where(tz >= P["TempSnow"], prec, 0.0)
            melt_s = (tz - P["TempMelt"]) * seas * (1 + 0.01 * rain_s) * dt                  # :162
            ice_s_ = (tavg if i < 2 else tz) * 7.0 * dt * summer                             # :164-168
            melt_s = np.maximum(np.minimum(melt_s + ice_s_, self.SnowCoverS[i]), 0.0)        # :170
            self.SnowCoverS[i] = self.SnowCoverS[i] + snow_s - melt_s
            snow += snow_s
            rain += rain_s
            melt += melt_s
            cover += self.SnowCoverS[i]
        snow /= 3
        rain /= 3
        melt /= 3
        cover /= 3
        self.TotalPrecipitation = self.TotalPrecipitation + (snow + rain)                    # :186
        rate = -(1 - P["Afrost"]) * self.FrostIndex - tavg * np.exp(-0.04 * P["Kfrost"] * cover / P["SnowWaterEquivalent"])
        fi = np.maximum(self.FrostIndex + rate * dt, 0)                                      # frost.py:68
        self.FrostIndex = np.where(fi > 57.0, 57.0, fi)                                      # :71
        o.update(Rain=rain, Snow=snow, SnowMelt=melt, SnowCover=cover, FrostIndex=self.FrostIndex.copy(),
                 isFrozenSoil=self.FrostIndex > P["FrostIndexThreshold"], SnowCoverS=np.stack(self.SnowCoverS),
                 TotalPrecipitation=self.TotalPrecipitation.copy())
        return o


def lai_term(kgb, lai):
    """leafarea.py:90."""
    return np.exp(-np.asarray(kgb, np.float64) * np.asarray(lai, np.float64))


def cf_pack(values, bits=16):
    """scale_factor / add_offset and the packed integers of a map, the CF way (NetCDF Users Guide, "Packed data values":
    add_offset = (max + min) / 2, scale_factor = (max - min) / (2^bits - 2)); the inverse of cf_unpack up to rounding."""
    v = np.asarray(values, np.float64)
    lo, hi = float(v.min()), float(v.max())
    scale = (hi - lo) / (2 ** bits - 2) if hi > lo else 1.0
    offset = (hi + lo) / 2
    return np.rint((v - offset) / scale).astype(np.int16), scale, offset


def cf_unpack(raw, scale_factor, add_offset, decode="float32"):
    """What the reference's reader hands on for a packed variable (xarray's CF decoding behind netcdf.py:231-232):
    raw * scale_factor + add_offset in the decoder's float type -- float32 for int16 data unless the attributes are
    float64."""
    t = np.float32 if decode == "float32" else np.float64
    return np.asarray(raw).astype(t) * t(scale_factor) + t(add_offset)
