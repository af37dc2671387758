"""Anchor grid [37,37,A,4] — lib/fpn/generate_anchors.py:39-126 of the reference (ratio
enumeration WITHOUT rounding, :110-111), float64 numpy, computed once at module construction."""
import numpy as np

from config import IM_SCALE


def _whctrs(a):
    w = a[2] - a[0] + 1
    h = a[3] - a[1] + 1
    return w, h, a[0] + 0.5 * (w - 1), a[1] + 0.5 * (h - 1)


def _mkanchors(ws, hs, x_ctr, y_ctr):
    ws, hs = ws[:, None], hs[:, None]
    return np.hstack((x_ctr - 0.5 * (ws - 1), y_ctr - 0.5 * (hs - 1), x_ctr + 0.5 * (ws - 1), y_ctr + 0.5 * (hs - 1)))


def generate_base_anchors(base_size=16, ratios=(0.5, 1, 2), scales=2 ** np.arange(3, 6)):
    ratios, scales = np.asarray(ratios), np.asarray(scales)
    w, h, xc, yc = _whctrs(np.array([1, 1, base_size, base_size]) - 1)
    ws = np.sqrt(w * h / ratios)          # no rounding
    ratio_anchors = _mkanchors(ws, ws * ratios, xc, yc)
    rows = []
    for ra in ratio_anchors:
        w, h, xc, yc = _whctrs(ra)
        rows.append(_mkanchors(w * scales, h * scales, xc, yc))
    return np.vstack(rows)


def generate_anchors(base_size=16, feat_stride=16, anchor_scales=(8, 16, 32), anchor_ratios=(0.5, 1, 2)):
    anchors = generate_base_anchors(base_size=base_size, ratios=anchor_ratios, scales=anchor_scales)
    shift = np.arange(0, IM_SCALE // feat_stride) * feat_stride
    sx, sy = np.meshgrid(shift, shift)
    shifts = np.stack([sx, sy, sx, sy], -1)
    return shifts[:, :, None] + anchors[None, None]
