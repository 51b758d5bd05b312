"""GPU x2 (NCCL): batch-sharded gradients == the 1-GPU global-batch gradients on CrossAttnRNN210 at the default dims.
Needs two visible GPUs (``gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp_nccl.py``); skipped on a 1-GPU box."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


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
                red = GradReducer(m, hooks=True)                 # broadcasts rank 0's parameters
                mine = shard_batch(full, rank, world)
                for p in m.parameters():
                    p.grad = None
                if mode == "hooks":
                    torch.manual_seed(55)
                    loss = m.training_step(mine, 0)
                    loss.backward()
                    red.finish()
                else:
                    torch.manual_seed(54)
                    step = GraphedTrainStep(m, mine, reducer=red)
                    torch.manual_seed(55)
                    step(mine)
                torch.cuda.synchronize()
                sharded = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
                red.remove()
                for p in m.parameters():
                    p.grad = None
                # the 1-GPU global-batch gradient, computed on every rank from the (now identical) parameters
                torch.manual_seed(55)
                loss = m.training_step(full, 0)
                loss.backward()
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
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


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
    out = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, res in out:
        print(rank, res)
        for (mode, precision), worst in res.items():
            # fp32: summation-order noise of a sum over 16 rows computed as two sums over 8; tensor-core mode: the
            # tf32/bf16 roundings differ between the sharded and the global products
            assert worst < (2e-5 if precision == "fp32" else 2e-2), (rank, mode, precision, worst)
