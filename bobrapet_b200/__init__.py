"""bobrapet_b200 — B200-native StoryRun DAG ready-frontier engine.

One hot path of bubustack/bobrapet (internal/controller/runs/dag.go: findReadySteps and
the per-iteration state machine around it), evaluated for a whole batch of StoryRuns by
hand-written sm_100a CUDA kernels behind the C ABI of include/bobrafrontier.h.
"""
from . import _abi  # noqa: F401
from ._abi import FrontierError, load  # noqa: F401
from .frontier import Frontier, FrontierGroup, TopologySet  # noqa: F401
from .records import make_layout, pack_state, unpack_result  # noqa: F401
