"""GPU x2 (NCCL): batch-sharded gradients == the 1-GPU global-batch gradients on CrossAttnRNN210 at the default dims.
Needs two visible GPUs (``gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp_nccl.py``); skipped on a 1-GPU box."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report_and_die(q, rank, text):
    """Hand the traceback to the parent, FLUSH the queue's feeder thread (os._exit right after put() loses the message
    and the parent waits for its whole timeout), then exit without waiting for the peer, which may be blocked in a
    collective."""
    q.put((rank, {"error": text}))
    q.close()
    q.join_thread()
    os._exit(1)


def _trace(rank, msg):
    """Progress marks of a worker (V2F_NCCL_TEST_TRACE=<dir>): where a rank is when a collective does not return."""
    d = os.environ.get("V2F_NCCL_TEST_TRACE")
    if d:
        with open(os.path.join(d, f"nccl_test_rank{rank}.log"), "a") as f:
            f.write(msg + "\n")


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import torch.nn as nn
    import torch.nn.functional as F
    import visuelle2_multimodal_fusion_b200.models.modules as mods
    import visuelle2_multimodal_fusion_b200.synth as synth
    from visuelle2_multimodal_fusion_b200.ddp import GradReducer, shard_batch
    from visuelle2_multimodal_fusion_b200.graphs import GraphedTrainStep
    from visuelle2_multimodal_fusion_b200.models import CrossAttnRNN210
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        cat_d, col_d, fab_d = synth.label_dicts()
        mods.resnet101_trunk = lambda: nn.Identity()
        torch.manual_seed(100 + rank)              # DIFFERENT init per rank: the reducer's broadcast must fix it
        m = CrossAttnRNN210.CrossAttnRNN(512, 512, 512, cat_d, col_d, fab_d, synth.STORE_N, 3).cuda().eval()
        m.on_train_epoch_start()
        B = 16
        data, feat = synth.make_batch(B, out_len=10, seed=9, feat_hw=10)
        full = (tuple(t.cuda() for t in data), feat.cuda())
        results = {}
        for mode in ("hooks", "graph"):
            for precision in ("fp32", "bf16"):
                m.precision = precision
                _trace(rank, f"{mode} {precision}: reducer")
                red = GradReducer(m, hooks=True)                 # broadcasts rank 0's parameters
                mine = shard_batch(full, rank, world)
                for p in m.parameters():
                    p.grad = None
                if mode == "hooks":
                    torch.manual_seed(55)
                    loss = m.training_step(mine, 0)
                    loss.backward()
                    red.finish()
                    del loss
                else:
                    torch.manual_seed(54)
                    _trace(rank, f"{mode} {precision}: capture")
                    step = GraphedTrainStep(m, mine, reducer=red)
                    torch.manual_seed(55)
                    _trace(rank, f"{mode} {precision}: replay")
                    step(mine)
                torch.cuda.synchronize()
                _trace(rank, f"{mode} {precision}: sharded step done")
                sharded = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
                if mode == "graph":
                    # a captured graph holds the communicator's kernels: it has to die before the process group does
                    # (destroy_process_group with the graph alive never returned)
                    step.close()
                    del step
                red.remove()
                for p in m.parameters():
                    p.grad = None
                # the 1-GPU global-batch gradient, computed on every rank from the (now identical) parameters
                torch.manual_seed(55)
                loss = m.training_step(full, 0)
                loss.backward()
                # a live loss keeps its autograd graph -- and the AccumulateGrad nodes, created on THIS stream -- alive;
                # the next GraphedTrainStep captures on its own stream and must get fresh ones
                del loss
                worst = 0.0
                for k, p in m.named_parameters():
                    if p.grad is None:
                        assert k not in sharded, k
                        continue
                    g = p.grad
                    d = float((sharded[k] - g).abs().max())
                    s = float(g.abs().max())
                    floor = 1e-6 if k.endswith("attn_linear.bias") else 1e-7 * (s + 1e-3)
                    worst = max(worst, (d - floor) / max(s, 1e-30))
                results[(mode, precision)] = worst
                _trace(rank, f"{mode} {precision}: compared, worst {worst:.3e}")
        q.put((rank, results))
    except Exception:
        import traceback
        _report_and_die(q, rank, traceback.format_exc())
    finally:
        dist.destroy_process_group()


def _worker_gtm(rank, world, port, q):
    """GTM_Visuelle2 (BatchNorm1d with batch statistics in its fusion network, GTM_Visuelle2.py:158) and
    Proposed_model_v3: with ddp.sync_batchnorm1d, 2 ranks x 8 items == the 1-GPU batch of 16 items."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import visuelle2_multimodal_fusion_b200.synth as synth
    from helpers import _restore_gtm_trunk, gtm_product_ctor
    from oracle.refshim import zero_dropout
    from visuelle2_multimodal_fusion_b200.ddp import GradReducer, shard_batch, sync_batchnorm1d
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        cat_d, col_d, fab_d = synth.label_dicts()
        results = {}
        for variant in ("gtm", "v3"):
            torch.manual_seed(7)
            try:
                m = gtm_product_ctor(variant)(32, 64, 12, 4, 1, 1, 1, cat_d, col_d, fab_d, synth.STORE_N, 52, 3, 0)
            finally:
                _restore_gtm_trunk()
            m = zero_dropout(m.cuda()).train()
            B = 16
            data, feat = synth.make_batch(B, demand=True, seed=5, feat_hw=3)
            data = (data[0][:, :12].contiguous(),) + data[1:]
            full = (tuple(t.cuda() for t in data), feat.cuda())
            red = GradReducer(m, hooks=True)
            state0 = {k: v.clone() for k, v in m.state_dict().items()}
            # 1-GPU global batch, per-rank statistics == global statistics
            loss = m.training_step(full, 0)
            loss.backward()
            red.finish()                                        # averages identical gradients: no-op
            ref = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
            ref_stats = {k: v.clone() for k, v in m.state_dict().items() if "running" in k}
            m.load_state_dict(state0)
            for p in m.parameters():
                p.grad = None
            assert sync_batchnorm1d(m) >= 1
            loss = m.training_step(shard_batch(full, rank, world), 0)
            loss.backward()
            red.finish()
            worst, detail = 0.0, []
            # absolute floor: a bias in front of BatchNorm has an exactly-zero gradient (2e-7 of summation noise in both
            # runs): differences are judged against the largest gradient of the model, not against that noise
            gmax = max(float(v.abs().max()) for v in ref.values())
            for k, p in m.named_parameters():
                if k not in ref:
                    continue
                d, sc = float((p.grad - ref[k]).abs().max()), float(ref[k].abs().max())
                detail.append(((d - 2e-7 - 1e-6 * gmax) / max(sc, 1e-6), k, d, sc))
            for k, v in m.state_dict().items():
                if "running" in k:
                    d, sc = float((v - ref_stats[k]).abs().max()), float(ref_stats[k].abs().max())
                    detail.append((d / max(sc, 1e-6), k, d, sc))
            detail.sort(reverse=True)
            worst, detail = detail[0][0], detail[:6]
            results[variant] = worst
            results[variant + ":detail"] = detail
            red.remove()
        q.put((rank, results))
    except Exception:
        import traceback
        _report_and_die(q, rank, traceback.format_exc())
    finally:
        dist.destroy_process_group()


def _collect(q, procs, timeout=150):
    """One result per worker; the first worker error fails the test at once (its peer is killed, not waited for)."""
    import queue
    out = []
    try:
        for _ in procs:
            try:
                rank, res = q.get(timeout=timeout)
            except queue.Empty:
                raise AssertionError(f"no result from the workers within {timeout} s (V2F_NCCL_TEST_TRACE=<dir> shows "
                                     "where each rank is)") from None
            assert "error" not in res, f"rank {rank}:\n{res['error']}"
            out.append((rank, res))
        for p in procs:
            p.join(60)
            assert p.exitcode == 0
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    return out


def test_two_gpu_sync_batchnorm1d_equals_global_batch():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + 7) % 2000
    procs = [ctx.Process(target=_worker_gtm, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = _collect(q, procs)
    for rank, res in out:
        print(rank, res)
        for variant, worst in res.items():
            if not variant.endswith(":detail"):
                assert worst < 1e-4, (rank, variant, worst, res[variant + ":detail"])


def test_two_gpu_sharded_gradients_equal_global_batch_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = _collect(q, procs)
    for rank, res in out:
        print(rank, res)
        for (mode, precision), worst in res.items():
            # fp32: summation-order noise of a sum over 16 rows computed as two sums over 8; tensor-core mode: the
            # tf32/bf16 roundings differ between the sharded and the global products
            assert worst < (2e-5 if precision == "fp32" else 2e-2), (rank, mode, precision, worst)
