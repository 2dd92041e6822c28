"""structures -- HydroModule mirror (reference: src/lisflood/hydrological_modules/structures.py:43-61).

Reservoirs and lakes interrupt the flow paths of the kinematic wave: every pixel just upstream of a structure becomes a
pit of LddKinematic, and the unmodified network is kept as LddStructuresKinematic to connect the inflow and the outflow
point of each structure.  On the device the model is levelled on LddStructuresKinematic and the channel kernel applies
the cut itself (lf_model_set_structures); this host mirror provides the maps for the init chain
(lisflood_code_b200/Lisflood_initial.py::initialise) and for HotPathModel."""
import numpy as np

from . import HydroModule
from ..global_modules import ldd_ops


class structures(HydroModule):
    input_files_keys = {'all': []}
    module_name = 'Structures'

    def __init__(self, structures_variable):
        self.var = structures_variable

    def initial(self):
        v = self.var
        land = v.maskinfo.land_mask
        v.LddStructuresKinematic = np.asarray(v.LddKinematic, np.float64).copy()                 # :46
        if v.option('InitLisflood'):                                                             # :49: not in the pre-run
            return
        ldd = np.asarray(v.LddKinematic, np.float64)
        ds = ldd_ops.downstream_index(ldd, land)
        struct = np.asarray(v.IsStructureKinematic) != 0
        # downstream(ldd, x): the value of the downstream cell, a pit keeps its own (PCRaster), :51-56
        v.IsUpsOfStructureKinematicC = np.where(ds >= 0, struct[np.maximum(ds, 0)], struct)
        v.LddKinematic = ldd_ops.lddrepair_codes(np.where(v.IsUpsOfStructureKinematicC, 5.0, ldd), land)   # :59
