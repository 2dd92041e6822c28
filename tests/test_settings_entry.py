"""Settings-XML surface and HydroModule call protocol (CPU parts) + the lisf1.py entry on the GPU."""
import os

import numpy as np
import pytest

XML = """<?xml version="1.0" encoding="UTF-8"?>
<lfsettings>
  <lfoptions><setoption name="SplitRouting" choice="%d"/><setoption name="InitLisflood" choice="0"/></lfoptions>
  <lfuser>
    <textvar name="PathRoot" value="%s"/><textvar name="StepStart" value="1"/><textvar name="StepEnd" value="%d"/>
  </lfuser>
  <lfbinding>
    <textvar name="MaskMap" value="$(PathRoot)/mask.npy"/><textvar name="StateFile" value="$(PathRoot)/state.npz"/>
    <textvar name="ForcingFile" value="$(PathRoot)/forcing.npz"/><textvar name="DisOut" value="$(PathRoot)/dis.npy"/>
    <textvar name="StepStart" value="$(StepStart)"/><textvar name="StepEnd" value="$(StepEnd)"/>
  </lfbinding>
</lfsettings>
"""


def _write_case(tmp, split, steps=2):
    from lisflood_code_b200 import synthetic
    S = synthetic.full_stack(40, 36, seed=21, split_routing=split)
    np.save(tmp / "mask.npy", S["mask"])
    np.savez(tmp / "state.npz", **{k: v for k, v in S.items() if k != "mask"})
    Fs = [synthetic.forcing(S, t, 21) for t in range(steps)]
    np.savez(tmp / "forcing.npz", **{k: np.stack([F[k] for F in Fs]) for k in Fs[0]})
    xml = tmp / "settings.xml"
    xml.write_text(XML % (1 if split else 0, str(tmp), steps))
    return S, Fs, str(xml)


def test_settings_xml_parsing(tmp_path):
    from lisflood_code_b200.global_modules.settings import LisSettings
    S, Fs, xml = _write_case(tmp_path, True)
    st = LisSettings(xml, ["-q", "--nancheck"])
    assert st.options["SplitRouting"] is True and st.options["nonInit"] is True and st.options["wateruse"] is False
    assert st.binding["MaskMap"] == str(tmp_path / "mask.npy") and st.binding["StepEnd"] == "2"
    assert st.flags["quiet"] and st.flags["nancheck"] and not st.flags["loud"]
    assert LisSettings.instance() is st
    st.check_supported()
    st.options["wateruse"] = True
    with pytest.raises(NotImplementedError):
        st.check_supported()


def test_module_call_order_is_enforced():
    from lisflood_code_b200.hotpath import HotPathModel
    from lisflood_code_b200.hydrological_modules.opensealed import opensealed

    class Fake(HotPathModel):
        def __init__(self):
            self.__dict__["_soil_calls"] = []
            self.__dict__["ran"] = 0

        def soil(self):
            self.__dict__["ran"] += 1

    v = Fake()
    with pytest.raises(RuntimeError):
        opensealed(v).dynamic()          # before soilloop.dynamic_canopy
    for who in HotPathModel._SOIL_SEQUENCE:
        v._soil_stage_call(who)
    assert v.ran == 1
    v._require_soil_stage_done()
    with pytest.raises(RuntimeError):
        v._require_soil_stage_done()


@pytest.mark.gpu
@pytest.mark.parametrize("split", [False, True])
def test_lisf1_entry_matches_oracle(gpu_lib, oracle, tmp_path, split):
    import lisf1
    from oracle import lisf_oracle_model as om
    S, Fs, xml = _write_case(tmp_path, split, steps=3)
    assert lisf1.main(xml, "-v") == 0
    dis = np.load(tmp_path / "dis.npy")
    O = om.OracleModel(S)
    for t, F in enumerate(Fs):
        O.step(F)
        want = O.var.ChanQAvg
        assert np.max(np.abs(dis[t] - want) / np.maximum(np.abs(want), 1e-12)) < 1e-8, t


def test_map_api_roundtrip(tmp_path):
    from lisflood_code_b200.global_modules import add1
    rng = np.random.default_rng(0)
    excluded = rng.random((7, 9)) < 0.3
    mi = add1.MaskInfo(excluded)
    m2 = rng.random((7, 9))
    c = add1.compressArray(m2)
    assert c.shape == (mi.num_pixels,) and np.array_equal(c, m2[~excluded])
    d = add1.decompress(c)
    assert np.array_equal(d[~excluded], c) and np.all(d[excluded] == -9999.0)
    assert mi.in_zero().shape == (mi.num_pixels,) and add1.makenumpy(2.5)[0] == 2.5
    np.save(tmp_path / "m.npy", m2)
    binding = {"beta": "0.6", "SomeMap": str(tmp_path / "m.npy")}
    assert isinstance(add1.loadmap("beta", binding), float) and add1.loadmap("beta", binding) == 0.6
    assert np.array_equal(add1.loadmap("SomeMap", binding), c)
    nm = add1.NumpyModified(np.zeros((3, 4)), ["vegetation", "pixel"])
    assert nm.values is nm and nm.dims == ["vegetation", "pixel"] and nm[1:].dims == ["vegetation", "pixel"]


RAW_XML = """<?xml version="1.0" encoding="UTF-8"?>
<lfsettings>
  <lfoptions><setoption name="SplitRouting" choice="%d"/><setoption name="drainedIrrigation" choice="%d"/></lfoptions>
  <lfuser><textvar name="PathRoot" value="%s"/><textvar name="PathMaps" value="$(PathRoot)/maps"/></lfuser>
  <lfbinding>
    <textvar name="MaskMap" value="$(PathRoot)/mask.npy"/>
    <textvar name="ForcingFile" value="$(PathRoot)/forcing.npz"/><textvar name="DisOut" value="$(PathRoot)/dis.npy"/>
    <textvar name="DtSec" value="%r"/><textvar name="DtSecChannel" value="3600"/>
%s
  </lfbinding>
</lfsettings>
"""


def _write_raw_case(tmp, case):
    """Settings + one .npy per raw static input of a golden init case (numbers go into the XML as they are)."""
    from conftest import load_golden
    g = load_golden(case)
    (tmp / "maps").mkdir()
    np.save(tmp / "mask.npy", g["mask"])
    lines = []
    raw = {k[5:]: v for k, v in g.items() if k.startswith("raw__")}
    for k in ("Forest", "DirectRunoff", "Water", "Irrigation", "Rice", "Other"):
        raw[k + "Fraction"] = g["state__" + k + "Fraction"]
    for k, v in raw.items():
        if v.ndim == 0:
            lines.append('    <textvar name="%s" value="%r"/>' % (k, float(v)))
        else:
            np.save(tmp / "maps" / (k + ".npy"), v)
            lines.append('    <textvar name="%s" value="$(PathMaps)/%s.npy"/>' % (k, k))
    split = bool(g["SplitRouting"])
    xml = tmp / "settings.xml"
    xml.write_text(RAW_XML % (1 if split else 0, 1 if split else 0, str(tmp), float(g["DtSec"]), "\n".join(lines)))
    return g, str(xml)


@pytest.mark.parametrize("case", ["init_21x26_split_sound", "init_19x23_single_6h"])
def test_lisf1_derives_the_model_state_from_raw_bindings(tmp_path, case):
    """No StateFile: the settings bind the raw static inputs by the reference's names and the modules' initial() derive
    the model state -- identical to what the reference's own initial() produced (golden)."""
    import lisf1
    from lisflood_code_b200.global_modules.settings import LisSettings
    g, xml = _write_raw_case(tmp_path, case)
    S = lisf1.model_state(LisSettings(xml, ["-q"]))
    for key, want in g.items():
        if not key.startswith(("soil__", "routing__", "surfgw__")):
            continue
        name = key.split("__", 1)[1]
        if name in S and name != "downstruct" and want.dtype.kind == "f":
            assert np.array_equal(np.asarray(S[name], np.float64), want, equal_nan=True), name
    assert S["SplitRouting"] == bool(g["SplitRouting"]) and S["NoRoutSteps"] == int(round(float(g["DtSec"]) / 3600.0))
    assert lisf1.main(xml, "-q", "-i") == 0       # --initonly stops before the device is touched


@pytest.mark.gpu
def test_lisf1_entry_from_raw_bindings_matches_oracle(gpu_lib, oracle, tmp_path):
    import lisf1
    from lisflood_code_b200 import synthetic
    from lisflood_code_b200.global_modules.settings import LisSettings
    from oracle import lisf_oracle_model as om
    g, xml = _write_raw_case(tmp_path, "init_21x26_split_sound")
    S = lisf1.model_state(LisSettings(xml, ["-q"]))
    S["kgb"] = 0.75 * 0.72
    Fs = [synthetic.forcing(S, t, 5) for t in range(3)]
    np.savez(tmp_path / "forcing.npz", **{k: np.stack([F[k] for F in Fs]) for k in Fs[0]})
    assert lisf1.main(xml, "-v") == 0
    dis = np.load(tmp_path / "dis.npy")
    O = om.OracleModel(S)
    for t, F in enumerate(Fs):
        O.step(F)
        want = O.var.ChanQAvg
        assert np.max(np.abs(dis[t] - want) / np.maximum(np.abs(want), 1e-12)) < 1e-8, t
