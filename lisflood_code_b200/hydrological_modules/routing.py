"""routing -- HydroModule mirror (reference: src/lisflood/hydrological_modules/routing.py:435-706).

The reference calls routing.dynamic(NoRoutingExecuted) NoRoutSteps times per model step
(Lisflood_dynamic.py:176-180).  On the device the sub-steps are one space-time wavefront (lf_model_channel), so the
call with NoRoutingExecuted == 0 executes all of them and the post-loop bookkeeping (Lisflood_dynamic.py:194-229);
the remaining calls of the step only check the sequence."""
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
        pass

    def initialSecond(self):
        pass   # river_router is built by lf_model_create from LddKinematic

    def dynamic(self, NoRoutingExecuted):
        if NoRoutingExecuted != self._expected:
            raise RuntimeError("routing.dynamic called with sub-step %d, expected %d" % (NoRoutingExecuted, self._expected))
        if NoRoutingExecuted == 0:
            self.var.channel()
        self._expected = (NoRoutingExecuted + 1) % self.var.NoRoutSteps
