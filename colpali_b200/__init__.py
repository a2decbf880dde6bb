"""colpali_b200 -- B200-native late-interaction hot path behind the colpali_engine API surface.

Public names mirror the reference (illuin-tech/colpali):
  score_multi_vector / score_single_vector   <- BaseVisualRetrieverProcessor (utils/processing_utils.py)
  ColbertLoss / ColbertPairwiseCELoss / ...  <- colpali_engine.loss (loss/late_interaction_losses.py)
  BiEncoderLoss / BiPairwiseCELoss / ...     <- colpali_engine.loss (loss/bi_encoder_losses.py)
  get_similarity_maps_from_embeddings        <- colpali_engine.interpretability (similarity_map_utils.py)
  fused_head                                 <- custom_text_proj + norm + mask tail of every Col* model forward
"""

from . import exchange
from ._lib import ColpaliB200Error
from .bi import (
    BiEncoderLoss,
    BiEncoderModule,
    BiNegativeCELoss,
    BiPairedEncoderLoss,
    BiPairwiseCELoss,
    BiPairwiseNegativeCELoss,
    BiSigmoidLoss,
    dense_dot,
    get_similarity_maps_from_embeddings,
    score_single_vector,
)
from .head import fused_head
from .install import install, uninstall
from .losses import (
    ColbertLoss,
    ColbertModule,
    ColbertNegativeCELoss,
    ColbertPairwiseCELoss,
    ColbertPairwiseNegativeCELoss,
    ColbertSigmoidLoss,
)
from .scoring import DocBank, QueryBlock, maxsim, score_multi_vector

__all__ = [
    "BiEncoderLoss",
    "BiEncoderModule",
    "BiNegativeCELoss",
    "BiPairedEncoderLoss",
    "BiPairwiseCELoss",
    "BiPairwiseNegativeCELoss",
    "BiSigmoidLoss",
    "dense_dot",
    "get_similarity_maps_from_embeddings",
    "ColbertLoss",
    "ColbertModule",
    "ColbertNegativeCELoss",
    "ColbertPairwiseCELoss",
    "ColbertPairwiseNegativeCELoss",
    "ColbertSigmoidLoss",
    "ColpaliB200Error",
    "DocBank",
    "exchange",
    "fused_head",
    "install",
    "uninstall",
    "QueryBlock",
    "maxsim",
    "score_multi_vector",
    "score_single_vector",
]
