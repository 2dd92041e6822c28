"""The product's settings-XML parser against the reference's OWN parser (global_modules/settings.py:502-607) on the
settings files the reference ships: every binding after $(var) substitution, every option the hot path reads, the
command-line flags.  Only where /root/reference exists (the build container)."""
import glob
import os

import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


@pytest.fixture(autouse=True)
def _keep_the_settings_singleton():
    """LisSettings(...) registers itself as the run's settings object: put back what was there."""
    from lisflood_code_b200.global_modules.settings import LisSettings
    before = LisSettings._instance
    yield
    LisSettings._instance = before


def _shipped_settings():
    if not ref_loader.available():
        return []
    root = os.path.normpath(os.path.join(ref_loader._R, "..", ".."))
    files = sorted(glob.glob(os.path.join(root, "tests", "data", "*", "settings", "*.xml")))
    files += sorted(glob.glob(os.path.join(root, "tests", "data", "*", "*.xml")))       # the lat / lon use case
    files.append(os.path.join(root, "src", "lisfloodSettings_reference.xml"))
    return [f for f in files if os.path.exists(f)]


@pytest.mark.parametrize("path", _shipped_settings(), ids=lambda p: "/".join(p.split(os.sep)[-3:]))
def test_bindings_and_options_equal_the_reference_parser(path):
    from lisflood_code_b200.global_modules.settings import LisSettings
    from oracle import ref_settings
    user, binding, options, _ = ref_settings.parse(path)
    mine = LisSettings(path)
    assert set(mine.binding) == set(binding)
    ref_project, my_project = user["ProjectDir"], mine.user["ProjectDir"]     # $(ProjectDir): the package directory differs
    for k, want in binding.items():
        got = mine.binding[k]
        if got != want:
            assert want.startswith(ref_project) and got == my_project + want[len(ref_project):], (k, want, got)
    assert len(binding) > 300
    for k, v in mine.options.items():
        assert options[k] == v, k
    assert {k: v for k, v in user.items() if k not in ("ProjectDir", "ProjectPath")} == \
        {k: v for k, v in mine.user.items() if k not in ("ProjectDir", "ProjectPath")}


@pytest.mark.parametrize("args", [[], ["-q"], ["-v", "-n"], ["--loud", "--initonly"], ["-qvlchtdnis"]])
def test_flags_equal_the_reference_parser(args):
    from lisflood_code_b200.global_modules.settings import LisSettings
    from oracle import ref_settings
    ref_settings.load()
    st = ref_settings.load()
    want = dict(st.LisSettings._flags.__wrapped__(args)) if hasattr(st.LisSettings._flags, "__wrapped__") else \
        dict(st.LisSettings._flags(list(args)))
    assert LisSettings._flags(args) == want
