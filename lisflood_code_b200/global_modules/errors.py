"""Warning / error types with the reference's names (reference: src/lisflood/global_modules/errors.py)."""


class LisfloodWarning(Warning):
    """Raised through warnings.warn, e.g. once per router on non-finite discharge
    (reference: hydrological_modules/kinematic_wave_parallel.py:180-184)."""

    def __init__(self, msg):
        self._msg = "\n\n ========================== LISFLOOD Warning =============================\n" + str(msg)
        super().__init__(self._msg)

    def __str__(self):
        return self._msg


class LisfloodError(Exception):
    def __init__(self, msg):
        self._msg = "\n\n ========================== LISFLOOD ERROR =============================\n" + str(msg)
        super().__init__(self._msg)

    def __str__(self):
        return self._msg
