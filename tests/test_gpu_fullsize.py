"""GPU: the kernels the benchmark times (persistent decoder, streaming attention, tcgen05 GEMMs, persistent GRU)
against the UNMODIFIED reference at its default dims E=A=H=512, Li=100, Lt=52 (train_dl.py:197-199), B=8, T=10/12/1.
tests/golden/full_*.pt come from oracle/make_golden_full.py (reference run in the build container): outputs,
attention maps and loss in full, every gradient by norm and by a 4096-value strided sample; the weights are
re-created from the fixture's seed and checked against the reference's per-tensor checksums."""
import pytest

from helpers import full_compare, full_model, full_run, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("kind", ["rnn210", "demand", "rnn21"])
def test_full_size_kernels_match_reference(kind, precision, tol):
    blob = load_golden("full_" + kind)
    m = full_model(blob, "cuda").eval()
    m.precision = precision
    out, loss, extras, grads, gfeat = full_run(m, blob, "cuda")
    n = full_compare(blob, out, loss, extras, grads, gfeat, tol)
    assert n >= 40
