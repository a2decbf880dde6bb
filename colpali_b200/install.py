"""Drop the B200 hot path into an importable ``colpali_engine`` (the reference) by attribute patching.

The reference has no plugin registry; its seams are plain Python attributes:

* the staticmethod ``BaseVisualRetrieverProcessor.score_multi_vector`` that every ``Col*Processor.score`` forwards
  to (colpali_engine/utils/processing_utils.py:132, e.g. models/qwen2/colqwen2/processing_colqwen2.py:115-125);
* the loss classes referenced by dotted path from training configs (scripts/configs/qwen2/train_colqwen2_model.yaml:25)
  and re-exported by ``colpali_engine.loss`` (loss/__init__.py:1-16), late-interaction and bi-encoder alike;
* the module-level ``get_similarity_maps_from_embeddings`` (interpretability/similarity_map_utils.py:9).

``install()`` swaps those attributes; ``uninstall()`` restores the originals.  Model heads are patched per model
file (three lines) because each ``forward`` owns its backbone call.
"""

from __future__ import annotations

import importlib
from typing import Dict, Tuple

_saved: Dict[Tuple[str, str], object] = {}

# the five concrete losses; the reference's ``ColbertModule`` base (helper methods only) is left alone
_LOSS_NAMES = ("ColbertLoss", "ColbertPairwiseCELoss", "ColbertNegativeCELoss",
               "ColbertPairwiseNegativeCELoss", "ColbertSigmoidLoss")
# the six concrete bi-encoder losses (loss/bi_encoder_losses.py:64-418)
_BI_LOSS_NAMES = ("BiEncoderLoss", "BiPairedEncoderLoss", "BiNegativeCELoss", "BiPairwiseCELoss",
                  "BiPairwiseNegativeCELoss", "BiSigmoidLoss")


def _swap(obj, name: str, new) -> None:
    key = (f"{getattr(obj, '__module__', '')}.{getattr(obj, '__qualname__', getattr(obj, '__name__', repr(obj)))}", name)
    if key not in _saved:
        _saved[key] = (obj, obj.__dict__.get(name, getattr(obj, name)))
    setattr(obj, name, new)


def install(scorer: bool = True, losses: bool = True, single_vector: bool = True, similarity_maps: bool = True) -> None:
    """Patch ``colpali_engine`` in this process.  Raises ImportError if the reference is not importable.

    ``single_vector`` also replaces ``score_single_vector`` (the Bi* processors call it with hidden-size embeddings,
    e.g. 1536-dim fp32): the dense kernel takes any dim and keeps fp32 operands in fp32.  ``similarity_maps`` replaces
    the module-level ``get_similarity_maps_from_embeddings`` (the Idefics3-ordering method of the processors,
    processing_utils.py:447, is left alone)."""
    from . import bi as _bi
    from . import losses as _losses
    from . import scoring as _scoring

    if scorer:
        pu = importlib.import_module("colpali_engine.utils.processing_utils")
        _swap(pu.BaseVisualRetrieverProcessor, "score_multi_vector", staticmethod(_scoring.score_multi_vector))
        if single_vector:
            _swap(pu.BaseVisualRetrieverProcessor, "score_single_vector", staticmethod(_bi.score_single_vector))
    if losses:
        for modname, names, src in (("colpali_engine.loss.late_interaction_losses", _LOSS_NAMES, _losses),
                                    ("colpali_engine.loss.bi_encoder_losses", _BI_LOSS_NAMES, _bi),
                                    ("colpali_engine.loss", _LOSS_NAMES, _losses),
                                    ("colpali_engine.loss", _BI_LOSS_NAMES, _bi)):
            try:
                mod = importlib.import_module(modname)
            except ImportError:
                if src is _losses:
                    raise
                continue  # a reference without bi_encoder_losses.py
            for n in names:
                if hasattr(mod, n):
                    _swap(mod, n, getattr(src, n))
    if similarity_maps:
        for modname in ("colpali_engine.interpretability.similarity_map_utils", "colpali_engine.interpretability"):
            try:
                mod = importlib.import_module(modname)
            except ImportError:  # the package's plotting half needs matplotlib / seaborn
                continue
            if hasattr(mod, "get_similarity_maps_from_embeddings"):
                _swap(mod, "get_similarity_maps_from_embeddings", _bi.get_similarity_maps_from_embeddings)


def uninstall() -> None:
    for (_, name), (obj, old) in list(_saved.items()):
        setattr(obj, name, old)
    _saved.clear()
