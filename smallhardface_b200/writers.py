"""Detection writers -- the on-disk result formats of the reference's datasets.

``write_detections`` reproduces ``lib/datasets/wider.py:143-170`` and ``lib/datasets/general.py:44-69`` byte for byte: per
image one text file ``<output_dir>/<image dir>/<image name with the extension replaced by txt>`` holding the image path,
the number of detections, then one ``x y w h score`` row per detection formatted ``'%d %d %d %d %g \\n'`` with
``int()`` truncation of the corner coordinates (note: width / height are differences of the TRUNCATED corners, and every
row ends with a space before the newline).  This is what the WIDER FACE evaluation toolbox
(``lib/wider_eval_tools/wider_eval.py:77-222``) and the official servers consume.
"""
from __future__ import annotations

import os
from typing import Sequence

import numpy as np


def format_detection_file(image_path: str, dets: np.ndarray) -> str:
    out = [image_path + "\n", str(len(dets)) + "\n"]
    for det in dets:
        out.append("%d %d %d %d %g \n" % (int(det[0]), int(det[1]), int(det[2]) - int(det[0]), int(det[3]) - int(det[1]),
                                           det[4]))
    return "".join(out)


def write_detections(image_paths: Sequence[str], dets_per_image: Sequence[np.ndarray], output_dir: str = "./output/",
                     extension: str = "jpg", strip_leading_slash: bool = False) -> None:
    """``image_paths`` as the imdb lists them (relative for WIDER: ``0--Parade/0_Parade_..jpg``; absolute for the
    ``general_*`` loader, which strips the leading '/': pass ``strip_leading_slash=True``, ``general.py:52-53``)."""
    if len(image_paths) != len(dets_per_image):
        raise ValueError("one detection array per image expected")
    for img_path, dets in zip(image_paths, dets_per_image):
        img_name = os.path.basename(img_path)
        img_dir = img_path[:img_path.find(img_name) - 1]
        if strip_leading_slash and img_dir[:1] == "/":
            img_dir = img_dir[1:]
        res_dir = os.path.join(output_dir, img_dir)
        if not os.path.isdir(res_dir):
            os.makedirs(res_dir)
        with open(os.path.join(res_dir, img_name.replace(extension, "txt")), "w") as f:
            f.write(format_detection_file(img_path, np.asarray(dets)))
