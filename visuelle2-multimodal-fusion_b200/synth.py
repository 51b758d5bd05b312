"""Seeded synthetic VISUELLE2-shaped batches.

The dataset is not in the container (and there is no network), so benchmarks and parity tests
use batches with the layout ``dataset_fusion.py:200-203,65`` produces:
``((X[B,W,2], y[B,W,H], cat[B], col[B], fab[B], store[B], temporal[B,4], gtrends[B,3,52]), images)``
or, for the new-product demand task, ``((ts[B,12], cat, ..., gtrends), images)``.
"""
import torch

CAT_N, COL_N, FAB_N, STORE_N = 27, 10, 59, 125   # label-dict sizes; store_num hard-coded train_dl.py:140


def label_dicts():
    return ({i: i for i in range(CAT_N)}, {i: i for i in range(COL_N)}, {i: i for i in range(FAB_N)})


def _sales(gen, *shape, dense=False):
    """Sparse small counts already divided by 53 (the code never normalises sales itself).  ``dense``: counts of
    8..53 with no zeros -- targets whose WAPE denominator (sum |gt|) is of the size of the numerator, for the
    training-trajectory fixtures (WAPE of order 100 %, where a 0.1-point bound means something)."""
    if dense:
        return torch.randint(8, 54, shape, generator=gen).float() / 53.0
    k = torch.randint(1, 11, shape, generator=gen).float() / 53.0
    keep = (torch.rand(shape, generator=gen) >= 0.6).float()
    return k * keep


def make_batch(batch, *, out_len=10, demand=False, seed=21, image_hw=299, images=True,
               feat_hw=None, num_trends=3, trend_len=52, dense_sales=False):
    """Returns ``(data_tuple, images_or_feature_map)`` on CPU.

    ``feat_hw`` set -> second element is a backbone feature map ``[B,2048,feat_hw,feat_hw]``
    (head-only runs); otherwise images ``[B,3,image_hw,image_hw]`` (or None if ``images=False``).
    """
    g = torch.Generator().manual_seed(seed)
    cat = torch.randint(0, CAT_N, (batch,), generator=g)
    col = torch.randint(0, COL_N, (batch,), generator=g)
    fab = torch.randint(0, FAB_N, (batch,), generator=g)
    store = torch.randint(0, STORE_N + 1, (batch,), generator=g)
    temporal = torch.rand(batch, 4, generator=g) * 0.97 + 0.03
    gt = torch.rand(batch, num_trends, trend_len, generator=g)
    lo = gt.min(dim=2, keepdim=True).values
    hi = gt.max(dim=2, keepdim=True).values
    gt = (gt - lo) / (hi - lo)                     # per-series MinMax, dataset_fusion.py:148-160
    if demand:
        head = (_sales(g, batch, 12, dense=dense_sales),)
    else:
        windows = 12 - 2 - out_len + 1             # dataset_fusion.py:98
        head = (_sales(g, batch, windows, 2, dense=dense_sales), _sales(g, batch, windows, out_len, dense=dense_sales))
    if feat_hw is not None:
        img = torch.randn(batch, 2048, feat_hw, feat_hw, generator=g).abs() * 0.5   # post-ReLU-like
    elif images:
        img = torch.randn(batch, 3, image_hw, image_hw, generator=g)
    else:
        img = None
    return head + (cat, col, fab, store, temporal, gt), img
