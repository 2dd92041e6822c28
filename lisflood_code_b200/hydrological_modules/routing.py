"""routing -- HydroModule mirror (reference: src/lisflood/hydrological_modules/routing.py:435-706).

The reference calls routing.dynamic(NoRoutingExecuted) NoRoutSteps times per model step
(Lisflood_dynamic.py:176-180).  On the device the sub-steps are one space-time wavefront (lf_model_channel), so the
call with NoRoutingExecuted == 0 executes all of them and the post-loop bookkeeping (Lisflood_dynamic.py:194-229);
the remaining calls of the step only check the sequence."""
import numpy as np

from . import HydroModule


class routing(HydroModule):
    input_files_keys = {'all': ['beta', 'ChanLength', 'Ldd', 'Channels', 'ChanGrad', 'ChanGradMin', 'CalChanMan', 'ChanMan',
                                'ChanBottomWidth', 'ChanDepthThreshold', 'ChanSdXdY', 'TotalCrossSectionAreaInitValue',
                                'PrevDischarge'],
                        'SplitRouting': ['CrossSection2AreaInitValue', 'PrevSideflowInitValue', 'CalChanMan2']}
    module_name = 'Routing'

    def __init__(self, routing_variable):
        self.var = routing_variable
        self._expected = 0

    def initial(self):
        """Drainage networks, channel geometry, kinematic-wave alpha and the initial channel state (reference:
        hydrological_modules/routing.py:61-351).  `self.var` is an InitialVariables; maps are compressed 1-D arrays and
        the PCRaster ldd operators are the NumPy restatements of global_modules/ldd_ops.py.  On a device-resident
        HotPathModel everything is already in place."""
        v = self.var
        if not hasattr(v, "defsoil"):
            return
        from ..global_modules import ldd_ops
        load, zeros, land = v.loadmap, v.maskinfo.in_zero, v.maskinfo.land_mask
        n = v.num_pixel
        v.avgdis = zeros()
        v.Beta = load('beta')
        v.InvBeta = 1 / v.Beta
        v.ChanLength = load('ChanLength').astype(float)
        v.InvChanLength = 1 / v.ChanLength
        v.NoRoutSteps = int(np.maximum(1, round(v.DtSec / v.DtSecChannel, 0)))   # :72
        if v.option('InitLisflood'):
            v.NoRoutSteps = 1
        v.DtRouting = v.DtSec / v.NoRoutSteps
        v.InvDtRouting = 1 / v.DtRouting
        v.InvNoRoutSteps = 1 / float(v.NoRoutSteps)
        # ---- drainage networks (:90-170) ----
        v.Ldd = ldd_ops.lddrepair_codes(load('Ldd'), land)
        ds = ldd_ops.downstream_index(v.Ldd, land)
        v.UpArea = ldd_ops.accuflux(ds, v.PixelArea)
        v.InvUpArea = 1 / v.UpArea
        v.IsChannel = np.asarray(load('Channels'), np.float64) != 0
        v.IsChannelKinematic = v.IsChannel.copy()
        v.IsStructureKinematic = np.zeros(n, bool)
        v.LddToChan = ldd_ops.lddrepair_codes(np.where(v.IsChannel, 5.0, v.Ldd), land)
        v.LddKinematic = ldd_ops.lddmask_codes(v.Ldd, v.IsChannel, land)   # 0 (missing value) off the channels
        v.AtLastPointC = ds < 0   # pit(Ldd), :127,160-161
        lddC = v.LddKinematic
        dsk = ldd_ops.downstream_index(lddC, land)
        v.downstruct = np.where(dsk >= 0, dsk, 0).astype("int32")
        v.downstruct[(lddC == 5) | (lddC == 0)] = n
        v.Catchments = ldd_ops.catchment_of_pits(ds).astype(np.int32)
        CatchArea = np.bincount(v.Catchments, weights=v.PixelArea)[v.Catchments]
        v.InvCatchArea = 1 / CatchArea
        # ---- channel geometry (:184-210) ----
        v.ChanGrad = np.maximum(load('ChanGrad'), load('ChanGradMin'))
        v.CalChanMan = load('CalChanMan')
        v.ChanMan = v.CalChanMan * load('ChanMan')
        v.ChanBottomWidth = load('ChanBottomWidth')
        ChanDepthThreshold = load('ChanDepthThreshold')
        ChanSdXdY = load('ChanSdXdY')
        v.ChanUpperWidth = v.ChanBottomWidth + 2 * ChanSdXdY * ChanDepthThreshold
        v.TotalCrossSectionAreaBankFull = 0.5 * ChanDepthThreshold * (v.ChanUpperWidth + v.ChanBottomWidth)
        half_bankfull = 0.5 * v.TotalCrossSectionAreaBankFull
        init_area = load('TotalCrossSectionAreaInitValue')
        v.TotalCrossSectionArea = np.where(init_area == -9999, half_bankfull, init_area)
        if v.option('SplitRouting'):
            cs2 = load('CrossSection2AreaInitValue')
            v.CrossSection2Area = np.where(cs2 == -9999, zeros(), cs2)
            prev = load('PrevSideflowInitValue')
            v.Sideflow1Chan = np.where(prev == -9999, zeros(), prev)
        # ---- kinematic-wave alpha at half bankfull depth (:220-232) ----
        depth_alpha = np.where(v.IsChannel, 0.5 * ChanDepthThreshold, 0.0)
        v.ChanWettedPerimeterAlpha = v.ChanBottomWidth + 2 * np.sqrt(np.square(depth_alpha) + np.square(depth_alpha * ChanSdXdY))
        alp_term = (v.ChanMan / (np.sqrt(v.ChanGrad))) ** v.Beta
        v.AlpPow = 2.0 / 3.0 * v.Beta
        v.ChannelAlpha = (alp_term * (v.ChanWettedPerimeterAlpha ** v.AlpPow)).astype(float)
        v.InvChannelAlpha = 1 / v.ChannelAlpha
        # ---- initial volume and discharge (:238-246, 325-340) ----
        v.ChanM3 = v.TotalCrossSectionArea * v.ChanLength
        v.ChanIniM3 = v.ChanM3.copy()
        v.ChanM3Kin = v.ChanIniM3.copy().astype(float)
        v.ChanQKin = np.where(v.ChannelAlpha > 0, (v.TotalCrossSectionArea / v.ChannelAlpha) ** v.InvBeta, 0).astype(float)
        v.CumQ = zeros()
        prev_q = load('PrevDischarge')
        v.ChanQ = np.where(prev_q == -9999, v.ChanQKin, prev_q)
        for nm in ("DischargeM3Out", "TotalQInM3", "sumDis", "sumInWB"):
            setattr(v, nm, zeros())

    def initialSecond(self):
        """Second (floodplain) line of routing: alpha of the virtual channel and the split limits from the pre-run's
        average discharge (reference: hydrological_modules/routing.py:353-397).  The device router itself is built by
        lf_model_create from LddKinematic."""
        v = self.var
        if not hasattr(v, "defsoil"):
            return
        v.ChannelAlpha2 = None
        if not v.option('SplitRouting'):
            return
        from ..global_modules import ldd_ops
        load = v.loadmap
        ChanMan2 = (v.ChanMan / v.CalChanMan) * load('CalChanMan2')
        alp_term2 = (ChanMan2 / (np.sqrt(v.ChanGrad))) ** v.Beta
        v.ChannelAlpha2 = (alp_term2 * (v.ChanWettedPerimeterAlpha ** v.AlpPow)).astype(float)
        v.InvChannelAlpha2 = 1 / v.ChannelAlpha2
        if v.option('InitLisflood'):
            return
        v.QLimit = load('AvgDis') * load('QSplitMult')
        v.M3Limit = v.ChannelAlpha * v.ChanLength * (v.QLimit ** v.Beta)
        v.Chan2M3Start = v.ChannelAlpha2 * v.ChanLength * (v.QLimit ** v.Beta)
        dsk = ldd_ops.downstream_index(v.LddKinematic, v.maskinfo.land_mask)
        v.Chan2QStart = v.QLimit - ldd_ops.upstream_sum(dsk, v.QLimit)
        v.Chan2M3Kin = v.CrossSection2Area * v.ChanLength + v.Chan2M3Start
        v.ChanM3Kin = v.ChanM3 - v.Chan2M3Kin + v.Chan2M3Start
        v.ChanM3Kin = np.where((v.ChanM3Kin < 0.0) & (v.ChanM3Kin > -0.0000001), 0.0, v.ChanM3Kin)
        v.Chan2QKin = (v.Chan2M3Kin * v.InvChanLength * v.InvChannelAlpha2) ** (v.InvBeta)
        v.ChanQKin = (v.ChanM3Kin * v.InvChanLength * v.InvChannelAlpha) ** (v.InvBeta)

    def dynamic(self, NoRoutingExecuted):
        if NoRoutingExecuted != self._expected:
            raise RuntimeError("routing.dynamic called with sub-step %d, expected %d" % (NoRoutingExecuted, self._expected))
        if NoRoutingExecuted == 0:
            self.var.channel()
        self._expected = (NoRoutingExecuted + 1) % self.var.NoRoutSteps
