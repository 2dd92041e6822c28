"""Settings-XML surface of LISFLOOD kept at the hot-path boundary
(reference: src/lisflood/global_modules/settings.py:369-607, src/lisflood/main.py:164-226).

Same file structure: <lfsettings><lfoptions><setoption name choice/></lfoptions><lfuser><textvar name value/>
</lfuser><lfbinding><textvar name value/></lfbinding></lfsettings>, `$(Name)` substitution from <lfuser> into
<lfbinding>, and the same command-line flags.  Only what the hot path consults is interpreted; the
meteo / IO / reporting machinery of the reference is out of scope.
"""
import getopt
import os
import xml.dom.minidom

# options the hot path reads (reference defaults: global_modules/default_options.py)
DEFAULT_OPTIONS = {"InitLisflood": False, "SplitRouting": False, "dynamicWave": False, "simulateLakes": False,
                   "simulateReservoirs": False, "simulatePolders": False, "inflow": False, "TransLoss": False,
                   "openwaterevapo": False, "wateruse": False, "repMBTs": False, "drainedIrrigation": False,
                   "simulatePF": False, "cropsEPIC": False, "repStressDays": False}
UNSUPPORTED_ON = ("dynamicWave", "simulatePolders", "inflow", "TransLoss", "openwaterevapo", "wateruse", "cropsEPIC")
FLAG_TABLE = [("quiet", "q"), ("veryquiet", "v"), ("loud", "l"), ("checkfiles", "c"), ("noheader", "h"), ("printtime", "t"),
              ("debug", "d"), ("nancheck", "n"), ("initonly", "i"), ("skipvalreplace", "s")]


class LisSettings(object):
    """Per-run settings object; `LisSettings.instance()` returns the last one built (settings.py:85-121)."""
    _instance = None

    def __init__(self, settings_file, sys_args=()):
        dom = xml.dom.minidom.parse(settings_file)
        self.settings_path = os.path.abspath(settings_file)
        self.settings_dir = os.path.normpath(os.path.dirname(self.settings_path))
        self.flags = self._flags(sys_args)
        self.options = self._options(dom)
        self.user, self.binding = self._bindings(dom)
        LisSettings._instance = self

    @classmethod
    def instance(cls):
        if cls._instance is None:
            raise RuntimeError("LisSettings not initialised")
        return cls._instance

    @staticmethod
    def _flags(sys_args):
        flags = {name: False for name, _ in FLAG_TABLE}
        try:
            opts, _ = getopt.getopt(list(sys_args), "".join(s for _, s in FLAG_TABLE), [n for n, _ in FLAG_TABLE])
        except getopt.GetoptError as e:
            raise SystemExit("lisf1: %s" % e)
        for o, _ in opts:
            for name, short in FLAG_TABLE:
                if o in ("-" + short, "--" + name):
                    flags[name] = True
        return flags

    @staticmethod
    def _options(dom):
        options = dict(DEFAULT_OPTIONS)
        for el in dom.getElementsByTagName("lfoptions"):
            for s in el.getElementsByTagName("setoption"):
                options[s.attributes["name"].value] = bool(int(s.attributes["choice"].value))
        options["nonInit"] = not options["InitLisflood"]
        return options

    def _bindings(self, dom):
        project_dir = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))   # settings.py:46
        user = {"ProjectDir": project_dir, "ProjectPath": project_dir,
                "SettingsDir": self.settings_dir, "SettingsPath": self.settings_dir}
        for el in dom.getElementsByTagName("lfuser"):
            for t in el.getElementsByTagName("textvar"):
                user[t.attributes["name"].value] = str(t.attributes["value"].value)
        binding = {}
        for el in dom.getElementsByTagName("lfbinding"):
            for t in el.getElementsByTagName("textvar"):
                binding[t.attributes["name"].value] = str(t.attributes["value"].value)
        last = None          # the reference keeps the previous substitution when a variable is missing (settings.py:554-558)
        for k, expr in binding.items():
            guard = 0
            while "$(" in expr and guard < 100:
                a1 = expr.find("$(")
                a2 = expr.find(")")
                name = expr[a1 + 2:a2]
                if name in user:
                    last = user[name]
                else:
                    print('no ', name, 'for', binding[k], ' in lfuser defined')
                    if last is None:
                        raise KeyError("no %s for %s in lfuser defined" % (name, k))
                expr = expr.replace(expr[a1:a2 + 1], last)
                guard += 1
            binding[k] = expr
        if "CalendarConvention" in binding:
            binding["calendar_type"] = binding["CalendarConvention"]      # settings.py:562
        return user, binding

    def check_supported(self):
        on = [o for o in UNSUPPORTED_ON if self.options.get(o)]
        if on:
            raise NotImplementedError("options outside the B200 hot path are switched on: %s (SURVEY.md §2: out of scope)"
                                      % ", ".join(on))
