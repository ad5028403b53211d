"""The host-orchestration tests of tests/test_engine_fake.py once more with the CUDA library behind the
engine instead of the NumPy test double (`-m gpu`): exon-first goldens of the reference, the deepcopy
continuation, the alive-overflow fallback, exon-first with top_n < 5, the pipelined batch, passes in flight,
the fraction-near-threshold flag, the reference-named likelihood methods and the error paths."""
import pytest

from kir_graph_b200 import engine
from tests.helpers import golden_names

pytestmark = pytest.mark.gpu


class LoggingCuda(engine.CudaBackend):
    """CudaBackend that records the launcher names (the fallback test counts gk_rank launches)."""

    def __init__(self):
        super().__init__()
        self.log = []

    def launch(self, name, *args, **kw):
        self.log.append(name)
        super().launch(name, *args, **kw)


@pytest.fixture
def on_cuda(monkeypatch):
    import tests.test_engine_fake as tef
    monkeypatch.setattr(tef, "FakeBackend", LoggingCuda)
    return tef


@pytest.mark.parametrize("name", golden_names("exonfirst"))
def test_exon_first_goldens(on_cuda, name):
    on_cuda.test_exon_first_class_against_reference(name)


@pytest.mark.parametrize("name", golden_names("typing"))
def test_allele_typing_goldens(on_cuda, name):
    on_cuda.test_allele_typing_class_against_reference(name)


def test_deepcopy_continuation(on_cuda):
    on_cuda.test_deepcopy_continues_search()


def test_alive_overflow_fallback(on_cuda, monkeypatch):
    on_cuda.test_alive_overflow_falls_back_to_exact_grids(monkeypatch)


def test_exon_first_top_n_below_five(on_cuda):
    on_cuda.test_exon_first_with_top_n_below_five_follows_the_reference()


def test_pipelined_batch_and_its_fallback(on_cuda, monkeypatch):
    on_cuda.test_pipelined_batch_equals_stepwise_and_falls_back(monkeypatch)


def test_passes_in_flight(on_cuda):
    on_cuda.test_pass_pipeline_equals_serial_passes()


def test_fraction_near_threshold_flag(on_cuda):
    on_cuda.test_fraction_near_the_select_best_threshold_is_reported()


@pytest.mark.parametrize("name", ["syn_a6_cn2", "syn_a12_cn4_nocorr", "worked_example_corr"])
def test_reference_named_likelihood_methods(on_cuda, name):
    on_cuda.test_reference_named_likelihood_methods(name)


def test_errors_and_empty(on_cuda):
    on_cuda.test_errors_and_empty()
