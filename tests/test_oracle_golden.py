"""CPU: the oracle restatement vs fixtures produced by the unmodified reference (oracle/make_golden.py)."""
import pytest
import torch

from helpers import assert_close, load_golden, oracle_run

RNN_CASES = ["rnn210_small", "rnn210_notf", "rnn21_small", "demand_small", "demand_notf"]
TOL = 1e-5   # fp32 contract, SURVEY.md section 8d


@pytest.mark.parametrize("name", RNN_CASES)
def test_oracle_matches_reference_golden(name):
    blob = load_golden(name)
    out, loss, extras, P, feat = oracle_run(blob)
    assert_close(out, blob["out"], TOL, name + ":out")
    assert_close(loss, blob["loss"], TOL, name + ":loss")
    for k, v in extras.items():
        assert_close(v, blob[k], TOL, f"{name}:{k}")
    loss.backward()
    assert_close(feat.grad, blob["grad_feat"], TOL, name + ":grad_feat")
    for k, g in blob["grads"].items():
        if g is None:
            assert P[k].grad is None or float(P[k].grad.abs().max()) == 0.0, f"{name}: {k} should get no grad"
        else:
            assert P[k].grad is not None, f"{name}: {k} has no oracle grad"
            assert_close(P[k].grad, g, TOL, f"{name}:grad:{k}")
