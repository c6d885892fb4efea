"""Base anchor enumeration (host side, a few dozen numbers computed once per net).

Same results as ``lib/layers/generate_anchors.py:11-86`` (checked against the reference's output in
tests/test_host_logic.py via tests/golden/anchors.npz), written in closed form: every anchor is a
w x h window centred on the centre of the ``base_size`` cell, w = round(sqrt(area/ratio)) * scale,
h = round(w0 * ratio) * scale, optionally shifted by ``shifts * stride``.
"""
from __future__ import annotations

import numpy as np


def generate_anchors(base_size=16, ratios=(0.5, 1, 2), scales=(8, 16, 32), shifts=(0,), strides=(0,)):
    ctr = 0.5 * (base_size - 1)
    rows = []
    for r in np.asarray(ratios, dtype=np.float64).ravel():
        w0 = np.round(np.sqrt(base_size * base_size / r))
        h0 = np.round(w0 * r)
        for sc, st in zip(np.asarray(scales).ravel(), np.asarray(strides).ravel()):
            w, h = w0 * sc, h0 * sc
            box = np.array([ctr - 0.5 * (w - 1), ctr - 0.5 * (h - 1), ctr + 0.5 * (w - 1), ctr + 0.5 * (h - 1)])
            sh = np.asarray(shifts, dtype=np.float64).ravel() * st
            for dy in sh:                      # meshgrid(shift, shift) raveled row-major: x fastest
                for dx in sh:
                    rows.append(box + np.array([dx, dy, dx, dy]))
    return np.vstack(rows)
