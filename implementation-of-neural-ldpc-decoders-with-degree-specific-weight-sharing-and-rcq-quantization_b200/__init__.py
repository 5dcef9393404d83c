"""B200-native batched LDPC decoding engine: flooding min-sum / neural degree-weighted min-sum / RCQ.

Drop-in for the decode hot path of the reference's Python classes (same names, constructors and
return tuples); the arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
``include/ldpc_b200.h``.  Importing the package loads ``libldpc_b200.so`` and fails if it is missing.
"""
from . import _lib as _lib_mod

_lib_mod.load()  # fail loudly if the CUDA library has not been built

from .engine import Engine, LdpcError, PinnedBuffer, TannerGraph, awgn_llr, count_errors  # noqa: E402,F401
from .ldpc_decoder import BasicMinSumDecoder, LDPCCode, create_test_ldpc_code, simulate_awgn_channel  # noqa: E402,F401
from .neural_2d_decoder import Neural2DMinSumDecoder, Neural2DOffsetMinSumDecoder  # noqa: E402,F401
from .neural_minsum_decoder import NeuralMinSumDecoder, NeuralOffsetMinSumDecoder  # noqa: E402,F401
from .rcq_decoder import NonUniformQuantizer, RCQMinSumDecoder, WeightedRCQDecoder  # noqa: E402,F401
from .simulation_framework import (LDPSimulator, SimulationConfig, SimulationResult,  # noqa: E402,F401
                                   create_test_decoders)
from .training_framework import GradientExplosionAnalyzer, PosteriorJointTrainer, TrainingConfig  # noqa: E402,F401
from . import codes  # noqa: E402,F401
