"""Oracle: spatial / temporal augmentation of event tensors and their boxes, numpy.

Restates, for ONE sample (a sequence of L frames) and a GIVEN augmentation state:
  data/utils/augmentor.py:228-247 (_zoom_out_and_rescale_tensor), :310-330 (_zoom_in_and_rescale_tensor), :404-410 (_flip_tensor),
  :457-478 (__call__: h-flip, then zoom-in or zoom-out), data/genx_utils/sequence_base.py:208-225 (time_flip_data: frames reversed,
  channels reversed), and the label methods data/genx_utils/labels.py:371-411 (zoom_in_and_rescale_), :437-459
  (zoom_out_and_rescale_), :482-497 (scale_), :499-502 (flip_lr_), :67-69 (remove_flat_labels_).
torch.nn.functional.interpolate(mode='nearest-exact') is restated as src = min(floor((dst + 0.5f) * (float(in) / float(out))), in - 1)
in fp32 (ATen's nearest_exact_idx); the fixtures of tests/golden/augment_cases.npz (made by running the reference's augmentor) pin it.
Label arithmetic is fp32 with Python-double constants rounded to fp32 when they meet the tensor, as torch does.
TEST INFRASTRUCTURE — see oracle/__init__.py.
"""
import numpy as np

F = np.float32


def nearest_exact_index(out_size: int, in_size: int) -> np.ndarray:
    scale = F(in_size) / F(out_size)
    idx = np.floor((np.arange(out_size, dtype=F) + F(0.5)) * scale).astype(np.int64)
    return np.minimum(idx, in_size - 1)


def augment_ev_repr(ev: np.ndarray, state: dict, is_reversed: bool = False) -> np.ndarray:
    """ev uint8 [L, C, H, W] of one sample; state = AugmentationState.to_dict()-like
    {'h_flip': {'active'}, 'zoom_in': {'active','x0','y0','factor'}, 'zoom_out': {...}}."""
    L, C, H, W = ev.shape
    x = ev
    if is_reversed:                      # sequence_base.py:214-217
        x = x[::-1, ::-1]
    if state['h_flip']['active']:        # augmentor.py:406-408
        x = x[..., ::-1]
    zi, zo = state['zoom_in'], state['zoom_out']
    if zi['active'] and zi['factor'] != 1:
        wh, ww = int(H / zi['factor']), int(W / zi['factor'])
        canvas = x[..., zi['y0']:zi['y0'] + wh, zi['x0']:zi['x0'] + ww]
        iy, ix = nearest_exact_index(H, canvas.shape[-2]), nearest_exact_index(W, canvas.shape[-1])
        x = canvas[..., iy[:, None], ix[None, :]]
    elif zo['active'] and zo['factor'] != 1:
        wh, ww = int(H / zo['factor']), int(W / zo['factor'])
        iy, ix = nearest_exact_index(wh, H), nearest_exact_index(ww, W)
        out = np.zeros_like(x)
        out[..., zo['y0']:zo['y0'] + wh, zo['x0']:zo['x0'] + ww] = x[..., iy[:, None], ix[None, :]]
        x = out
    return np.ascontiguousarray(x)


def _scale(rows, mult, hw):
    """labels.py:482-497.  rows fp32 [n, >=5] (t, x, y, w, h, ...); returns (rows, new hw)."""
    ht, wd = mult * hw[0], mult * hw[1]
    if len(rows) == 0 or mult == 1:
        return rows, hw if mult == 1 else (ht, wd)
    m = F(mult)
    x1 = np.minimum((rows[:, 1] + rows[:, 3]) * m, F(wd - 1))
    y1 = np.minimum((rows[:, 2] + rows[:, 4]) * m, F(ht - 1))
    rows = rows.copy()
    rows[:, 1] = rows[:, 1] * m
    rows[:, 2] = rows[:, 2] * m
    rows[:, 3] = x1 - rows[:, 1]
    rows[:, 4] = y1 - rows[:, 2]
    return rows[(rows[:, 3] > 0) & (rows[:, 4] > 0)], (ht, wd)


def augment_labels(rows: np.ndarray, hw, state: dict) -> np.ndarray:
    """rows fp32 [n, 8] of one label frame -> transformed rows (flat boxes removed)."""
    rows = rows.astype(F).copy()
    H, W = hw
    if len(rows) == 0:
        return rows
    if state['h_flip']['active']:        # labels.py:499-502
        rows[:, 1] = F(W - 1) - rows[:, 1] - rows[:, 3]
    zi, zo = state['zoom_in'], state['zoom_out']
    if zi['active'] and zi['factor'] != 1:
        f, zx0, zy0 = zi['factor'], zi['x0'], zi['y0']
        zh, zw = H / f, W / f
        zx1, zy1 = min(zx0 + zw, W - 1), min(zy0 + zh, H - 1)
        cl = lambda v, lo, hi: np.minimum(np.maximum(v, F(lo)), F(hi))  # noqa: E731
        x0, y0 = cl(rows[:, 1], zx0, zx1 - 1), cl(rows[:, 2], zy0, zy1 - 1)
        x1, y1 = cl(rows[:, 1] + rows[:, 3], zx0, zx1 - 1), cl(rows[:, 2] + rows[:, 4], zy0, zy1 - 1)
        rows[:, 1], rows[:, 2], rows[:, 3], rows[:, 4] = x0 - F(zx0), y0 - F(zy0), x1 - x0, y1 - y0
        rows = rows[(rows[:, 3] > 0) & (rows[:, 4] > 0)]
        rows, _ = _scale(rows, f, (zh, zw))
    elif zo['active'] and zo['factor'] != 1:
        rows, _ = _scale(rows, 1 / zo['factor'], (H, W))
        if len(rows):
            rows[:, 1] = rows[:, 1] + F(zo['x0'])
            rows[:, 2] = rows[:, 2] + F(zo['y0'])
    return rows
