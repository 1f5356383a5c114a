"""B200-native guided-DDIM + exemplar-retrieval hot path of RAG-Gesture.

Importing the package registers the CUDA-backed classes under the reference's registry names
(`MotionDiffusion`, `ReGestureTransformer`, `EfficientSelfAttention`, `EfficientCrossAttention`,
`MSELoss`).  The shared library is loaded on first use; there is no CPU fallback."""
__version__ = "0.1.0"

from . import config, synthetic  # noqa: F401
from .mogen_api import (ARCHITECTURES, ATTENTIONS, LOSSES, MODELS, SUBMODULES, build_architecture,  # noqa: F401
                        build_attention, build_loss, build_submodule, register_into)
from .architecture import GuidedPipeline, MotionDiffusion, MSELoss  # noqa: F401
from .mogen_api import (DecoderLayer, EfficientCrossAttention, EfficientSelfAttention, FFN,  # noqa: F401
                        ReGestureTransformer, StylizationBlock)
from .diffusion import SpacedDiffusion, build_diffusion, get_named_beta_schedule, space_timesteps  # noqa: F401
from .retrieval import RetrievalDatabase, discourse_retrieval, map_conns_to_prominence  # noqa: F401
