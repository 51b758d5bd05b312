"""CPU: the oracle restatement vs fixtures produced by the unmodified reference (oracle/make_golden.py)."""
import pytest
import torch

from helpers import assert_close, load_golden, noise_only_grads, oracle_run

RNN_CASES = ["rnn210_small", "rnn210_notf", "rnn21_small", "demand_small", "demand_notf"]
TOL = 1e-5   # fp32 contract, SURVEY.md section 8d


GTM_CASES = ["gtm_demand_eval", "gtm_demand_train", "gtm_sofore1_train", "gtm_ar_eval", "v4_demand_train",
             "v4_sofore10_eval", "v3_demand_train", "v1_demand_train", "v2_demand_train", "m4ft_demand_train",
             "m4ft_sofore10_eval"]


@pytest.mark.parametrize("name", RNN_CASES + GTM_CASES)
def test_oracle_matches_reference_golden(name):
    # GTM family: two fp32 CPU evaluations of the same graph already differ by 1.1e-5 on one tensor
    # (conv / einsum summation order amplified by BatchNorm batch statistics), so the fp32 noise floor
    # there is taken as 3e-5 (1e-4 for gtm_sofore1_train, whose 30-row batch holds only 3 distinct items, so
    # the BatchNorm batch variance is tiny and rounding is amplified); the RNN family stays at 1e-5.  Gradients that are exactly zero
    # in exact arithmetic (a bias feeding train-mode BatchNorm) are rounding noise ~2e-6 on both sides.
    tol = TOL if name in RNN_CASES else (1e-4 if name == "gtm_sofore1_train" else 3e-5)
    _check(name, tol, 1e-7)


def _check(name, TOL, floor):
    blob = load_golden(name)
    out, loss, extras, P, feat = oracle_run(blob)
    assert_close(out, blob["out"], TOL, name + ":out")
    assert_close(loss, blob["loss"], TOL, name + ":loss")
    for k, v in extras.items():
        assert_close(v, blob[k], TOL, f"{name}:{k}")
    loss.backward()
    noisy = noise_only_grads(blob)
    assert_close(feat.grad, blob["grad_feat"], TOL, name + ":grad_feat")
    for k, g in blob["grads"].items():
        if g is None:
            assert P[k].grad is None or float(P[k].grad.abs().max()) == 0.0, f"{name}: {k} should get no grad"
        else:
            assert P[k].grad is not None, f"{name}: {k} has no oracle grad"
            assert_close(P[k].grad, g, TOL, f"{name}:grad:{k}", floor=2e-5 if k in noisy else floor)
