"""Host-side mirrors of the reference's hydrological modules that sit on the hot path
(reference: src/lisflood/hydrological_modules/)."""


class HydroModule(object):
    """Same protocol as the reference's HydroModule base (hydrological_modules/__init__.py:49-76):
    class attributes `input_files_keys`, `module_name`; methods initial()/dynamic()."""
    input_files_keys = None
    module_name = None

    @classmethod
    def check_input_files(cls, option):
        return True
