"""B200-native conditional-flow-matching decode for Matcha-TTS-24k (one hot path).

Public surface:
  CFM                      drop-in for ``matcha.models.components.flow_matching.CFM``
  install()                patch that symbol inside an importable ``matcha`` package
  ShardedCFM               one batch decoded across all GPUs of the box from one process (utterance-sharded, no collective)
  EstimatorWeights, EstimatorConfig, weight_spec
  native                   ctypes binding of the C-ABI library (``include/cfm_b200.h``)
"""
from . import synthetic  # noqa: F401
from .estimator import EstimatorConfig, EstimatorWeights, config_from_decoder_params, weight_spec  # noqa: F401
from .cfm import CFM, install, lengths_from_mask  # noqa: F401,E402
from .sharding import ShardedCFM, gather_outputs, shard_utterances  # noqa: F401,E402
from . import _native as native  # noqa: F401,E402
