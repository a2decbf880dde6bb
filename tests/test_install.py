"""CPU: install()/uninstall() swap exactly the reference's seams (exercised on a stub package that has the
reference's attribute layout, so it runs without the reference; tests/test_reference_replay_gpu.py does the same on the
real package from baseline/_ref)."""
import sys
import textwrap

import colpali_b200 as cb


def _make_stub(tmp_path):
    root = tmp_path / "colpali_engine"
    (root / "utils").mkdir(parents=True)
    (root / "loss").mkdir()
    (root / "__init__.py").write_text("")
    (root / "utils" / "__init__.py").write_text("")
    (root / "utils" / "processing_utils.py").write_text(textwrap.dedent("""
        class BaseVisualRetrieverProcessor:
            @staticmethod
            def score_multi_vector(qs, ps, batch_size=128, device=None):
                return "reference-multi"
            @staticmethod
            def score_single_vector(qs, ps, device=None):
                return "reference-single"
        class ColStubProcessor(BaseVisualRetrieverProcessor):
            def score(self, qs, ps, device=None, **kw):
                return self.score_multi_vector(qs, ps, device=device, **kw)
    """))
    (root / "loss" / "late_interaction_losses.py").write_text(textwrap.dedent("""
        class ColbertModule: pass
        class ColbertLoss(ColbertModule): pass
        class ColbertPairwiseCELoss(ColbertModule): pass
        class ColbertSigmoidLoss(ColbertModule): pass
    """))
    (root / "loss" / "bi_encoder_losses.py").write_text(textwrap.dedent("""
        class BiEncoderModule: pass
        class BiEncoderLoss(BiEncoderModule): pass
        class BiPairedEncoderLoss(BiEncoderModule): pass
        class BiSigmoidLoss(BiEncoderModule): pass
    """))
    (root / "loss" / "__init__.py").write_text(
        "from .late_interaction_losses import ColbertLoss, ColbertModule, ColbertPairwiseCELoss, ColbertSigmoidLoss\n"
        "from .bi_encoder_losses import BiEncoderLoss, BiEncoderModule, BiSigmoidLoss\n")
    (root / "interpretability").mkdir()
    (root / "interpretability" / "similarity_map_utils.py").write_text(
        "def get_similarity_maps_from_embeddings(*a, **k):\n    return 'reference-maps'\n")
    (root / "interpretability" / "__init__.py").write_text(
        "from .similarity_map_utils import get_similarity_maps_from_embeddings\n")


def test_install_and_uninstall(tmp_path, monkeypatch):
    _make_stub(tmp_path)
    monkeypatch.syspath_prepend(str(tmp_path))
    for k in [k for k in sys.modules if k.startswith("colpali_engine")]:
        monkeypatch.delitem(sys.modules, k)
    import colpali_engine.interpretability as I  # noqa: E741
    import colpali_engine.loss as L
    import colpali_engine.utils.processing_utils as pu

    proc = pu.ColStubProcessor()
    assert proc.score(1, 2) == "reference-multi"
    cb.install()
    try:
        assert pu.BaseVisualRetrieverProcessor.score_multi_vector is cb.score_multi_vector
        assert pu.ColStubProcessor.score_multi_vector is cb.score_multi_vector          # subclasses inherit the patch
        assert L.ColbertLoss is cb.ColbertLoss and L.ColbertPairwiseCELoss is cb.ColbertPairwiseCELoss
        assert L.late_interaction_losses.ColbertLoss is cb.ColbertLoss                   # dotted-path configs resolve to it
        assert L.ColbertSigmoidLoss is cb.ColbertSigmoidLoss
        assert L.ColbertModule is not cb.ColbertModule                                   # the helper base class stays the reference's
        # score_single_vector (Bi* processors, hidden-size fp32 embeddings): the dense fp32 kernel, any dim
        assert pu.BaseVisualRetrieverProcessor.score_single_vector is cb.score_single_vector
        # bi-encoder losses and the similarity-map helper
        assert L.BiEncoderLoss is cb.BiEncoderLoss and L.BiSigmoidLoss is cb.BiSigmoidLoss
        assert L.bi_encoder_losses.BiPairedEncoderLoss is cb.BiPairedEncoderLoss      # not re-exported by the package (loss/__init__.py)
        assert L.BiEncoderModule is not cb.BiEncoderModule
        assert I.get_similarity_maps_from_embeddings is cb.get_similarity_maps_from_embeddings
        assert I.similarity_map_utils.get_similarity_maps_from_embeddings is cb.get_similarity_maps_from_embeddings
        L.BiEncoderLoss(temperature=0.02, pos_aware_negative_filtering=False, max_batch_size=1024, filter_threshold=0.95,
                        filter_factor=0.5)
        # constructible with the reference's keyword arguments (scripts/configs/**/*.yaml)
        L.ColbertPairwiseCELoss(temperature=0.02, normalize_scores=True, use_smooth_max=False,
                                pos_aware_negative_filtering=False, max_batch_size=1024, tau=0.1, norm_tol=1e-3,
                                filter_threshold=0.95, filter_factor=0.5)
    finally:
        cb.uninstall()
    assert proc.score(1, 2) == "reference-multi"
    assert L.ColbertLoss.__module__.startswith("colpali_engine")
    assert L.BiEncoderLoss.__module__.startswith("colpali_engine")
    assert I.get_similarity_maps_from_embeddings() == "reference-maps"
    assert pu.BaseVisualRetrieverProcessor.score_single_vector(1, 2) == "reference-single"
