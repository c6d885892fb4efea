"""Image-pyramid detection driver on one GPU: the device-resident equivalent of ``lib/test.py:109-178``
(``detect``) + ``lib/test.py:21-106`` (``forward_net``).

Per image: one uint8 upload; then for each pyramid scale and (optionally) its mirror the level blob is
produced on the device (``shf_preprocess_level``), the net runs, and the pass's detections (un-mirrored,
unscaled, score > 0.05) are appended to a device-side list; finally box voting or NMS runs batched over
all images of the call.  Only the final boxes leave the GPU -- the reference copies a float32 blob per
level to the device, both head blobs back (``proposal_layer.py:96-98``), and runs decode, sort and the
O(n * clusters) ``bbox_vote`` loop in NumPy under the GIL.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np
import torch

from . import caffe_proto as cp
from . import lib as L
from .engine import GpuNet, _ptr, _stream
from .graph import NetSpec, TEST, load_weights


@dataclass
class DetectConfig:
    """The ``cfg.TEST.*`` / ``cfg.*`` keys the hot path reads (``configs/default.toml``)."""
    scales: Sequence[int] = (100, 300, 600, 1000, 1400)        # TEST.SCALES
    pyramid_base_size: Sequence[int] = (800, 1200)             # TEST.PYRAMID_BASE_SIZE
    max_size: int = 2000                                       # TEST.MAX_SIZE (single-scale mode)
    flip: bool = True                                          # TEST.FLIP
    nms_method: str = "BBOX_VOTE"                              # TEST.NMS_METHOD
    nms_thresh: float = 0.4                                    # TEST.NMS_THRESH
    thresh: float = 0.05                                       # test_net(thresh=0.05)
    n_dets_per_module: int = 10000                             # TEST.N_DETS_PER_MODULE
    score_thresh: float = 0.002                                # TEST.SCORE_THRESH
    anchor_min_size: float = 0.0                               # TEST.ANCHOR_MIN_SIZE
    max_resolution: int = 16                                   # MAX_RESOLUTION
    pixel_means: Sequence[float] = (102.9801, 115.9465, 122.7717)   # PIXEL_MEANS
    nms_mode: int = 0                                          # 0 = cpu_nms (>=), 1 = gpu_nms (>)
    # rows of the device result buffers are sized from the input (cap / 2 + 1 for voting, cap for NMS: nothing can be
    # truncated); ``gather_rows`` only sizes the multi-GPU all-gather payload, and exceeding it raises (parallel.py)
    gather_rows: int = 4096
    # not a reference key: pyramid levels with im_scale >= this run the convs on the fast f16+f8 operand format, smaller
    # (error-magnifying) levels on precise split fp16; None = precise everywhere (see GpuNet / tools/precision_model.py)
    fast_min_scale: float | None = 0.9


def compute_scaling_factor(im_shape, target_size, max_size):
    """``lib/utils/test_utils.py:8-26`` (host scalar logic)."""
    lo, hi = min(im_shape[0], im_shape[1]), max(im_shape[0], im_shape[1])
    s = float(target_size) / float(lo)
    if np.round(s * hi) > max_size:
        s = float(max_size) / float(hi)
    return s


def pyramid_scales(im_shape, cfg: DetectConfig):
    """``lib/test.py:131-137``; single-scale mode ``lib/test.py:119-122`` when one scale is configured."""
    if len(cfg.scales) > 1:
        base = compute_scaling_factor(im_shape, cfg.pyramid_base_size[0], cfg.pyramid_base_size[1])
        return [float(s) / cfg.pyramid_base_size[0] * base for s in cfg.scales]
    return [compute_scaling_factor(im_shape, cfg.scales[0], cfg.max_size)]


def level_geometry(h, w, s, mult=16):
    """Resized size (cv2: rint(src * f), half to even) and the x16-padded size (``lib/test.py:34-36``)."""
    oh = h if s == 1.0 else int(np.rint(h * s))
    ow = w if s == 1.0 else int(np.rint(w * s))
    return oh, ow, -(-oh // mult) * mult, -(-ow // mult) * mult


class Detector:
    def __init__(self, prototxt, caffemodel, device="cuda:0", cfg: DetectConfig | None = None):
        self.cfg = cfg or DetectConfig()
        self.device = torch.device(device)
        net_param = cp.read_net_text(prototxt) if isinstance(prototxt, str) else prototxt
        model = cp.read_net_binary(caffemodel) if isinstance(caffemodel, str) else caffemodel
        spec = NetSpec(net_param, TEST)
        params = load_weights(spec, model)
        self.net = GpuNet(spec, params, device, pre_nms_topn=self.cfg.n_dets_per_module,
                          score_thresh=self.cfg.score_thresh, min_size=self.cfg.anchor_min_size,
                          fast_min_scale=self.cfg.fast_min_scale)
        self._means = (C.c_double * 3)(*self.cfg.pixel_means)
        self._bufs = {}
        self._staging = {}                   # (slot, shape) -> [pinned tensor, event of its last upload]
        self._flip = 0
        self._post_stream = None             # box voting / NMS of call i overlaps the conv stack of call i+1
        self.max_batch_bytes = 40e9          # activation budget used to size per-level batches (180 GB HBM per GPU)

    # ------------------------------------------------------------------------------------------
    def upload(self, images: List[np.ndarray]) -> List[torch.Tensor]:
        """uint8 HWC BGR host images -> device tensors (through pinned staging, async on the current stream)."""
        out = []
        for i, im in enumerate(images):
            if im.dtype != np.uint8 or im.ndim != 3 or im.shape[2] != 3:
                raise ValueError("images must be uint8 HxWx3 (cv2.imread layout)")
            key = (i, im.shape)
            st = self._staging.get(key)
            if st is None:                   # page-locked staging is allocated once per (slot, shape) and reused
                if len(self._staging) >= 256:
                    self._staging.clear()
                st = self._staging[key] = [torch.empty(im.shape, dtype=torch.uint8, pin_memory=True), None]
            if st[1] is not None:
                st[1].synchronize()          # the previous upload from this buffer has been consumed
            st[0].numpy()[...] = im
            out.append(st[0].to(self.device, non_blocking=True))
            st[1] = torch.cuda.Event()
            st[1].record()
        return out

    def _buffers(self, batch: int, passes: int):
        # two buffer sets used alternately: the post-processing of one call (side stream, one CTA per image -- it would
        # leave 140 SMs idle on the main stream) may still be running when the next call starts filling the other set
        self._flip ^= 1
        key = (batch, passes, self._flip)
        b = self._bufs.get(key)
        if b is None:
            cap = passes * self.net.cfg["pre_nms_topn"]
            dev = self.device
            ws_bytes = int(L.load().shf_postprocess_workspace(batch, cap))
            # voting emits at most one row per cluster of >= 2 members plus one last singleton; NMS at most every row
            out_cap = cap // 2 + 1 if self.cfg.nms_method == "BBOX_VOTE" else cap
            b = dict(cap=cap, out_cap=out_cap, guard=torch.zeros_like(self.net.guard),
                     dets=torch.empty((batch, cap, 5), dtype=torch.float32, device=dev),
                     offs=torch.zeros((batch, passes + 1), dtype=torch.int32, device=dev),
                     seg_begin=(torch.arange(batch, dtype=torch.int32, device=dev) * cap).contiguous(),
                     seg_end=torch.empty((batch,), dtype=torch.int32, device=dev),
                     out_dets=torch.empty((batch, out_cap, 5), dtype=torch.float32, device=dev),
                     out_idx=torch.empty((batch, out_cap), dtype=torch.int32, device=dev),
                     out_count=torch.zeros((batch,), dtype=torch.int32, device=dev),
                     ws=torch.empty((ws_bytes,), dtype=torch.uint8, device=dev), ws_bytes=ws_bytes)
            self._bufs[key] = b
        return b

    def _level_batch(self, stacked: torch.Tensor, s: float, flips):
        """One pyramid level of several same-sized images (stacked (n, h, w, 3) uint8) and their mirrors as a single
        (n * len(flips), 3, HP, WP) batch, produced by one launch."""
        n, h, w = stacked.shape[0], stacked.shape[1], stacked.shape[2]
        oh, ow, hp, wp = level_geometry(h, w, s, self.cfg.max_resolution)
        data = torch.empty((n * len(flips), 3, hp, wp), dtype=torch.float32, device=self.device)
        L.call("shf_preprocess_level_batched", _ptr(stacked), n, h, w, _ptr(data), oh, ow, hp, wp, float(s), len(flips),
               self._means, _stream())
        self.net.launches += 1
        return data, (oh, ow, s)

    def detect_device(self, dev_images: List[torch.Tensor]):
        """Runs the whole pipeline for a batch of device-resident images; returns the device buffers
        (out_dets (B,max,5), out_idx (B,max), out_count (B,), dets (B,cap,5)) without synchronising.  The final
        vote / NMS launch is queued on a side stream: ``wait_results(b)`` (or ``download``) orders a consumer after it.
        A returned set stays valid until the second-next call with the same batch shape.

        Same-sized images are stacked per pyramid level together with their mirrored copies, so the conv stack
        sees N = images x flips per launch (the small levels would not fill 148 SMs otherwise); the
        ProposalLayer tail and the per-pass bookkeeping stay per image, in the reference's pass order
        (scale 0, scale 0 mirrored, scale 1, ... -- ``lib/test.py:141-155``)."""
        cfg = self.cfg
        flips = (False, True) if cfg.flip else (False,)
        nf = len(flips)
        nscales = len(cfg.scales) if len(cfg.scales) > 1 else 1
        passes = nscales * nf
        B = len(dev_images)
        b = self._buffers(B, passes)
        if b.get("done") is not None:
            torch.cuda.current_stream().wait_event(b["done"])      # this set's previous post-processing has finished
        b["offs"].zero_()
        b["guard"].zero_()
        self.net.guard = b["guard"]          # this call's range-guard slots (read back in download())
        b["images"] = dev_images
        groups = {}
        for i, img in enumerate(dev_images):
            groups.setdefault((img.shape[0], img.shape[1]), []).append(i)
        # buffer slot k holds image order[k]: same-sized images occupy consecutive slots, which is what the batched
        # tail kernels index by; download() undoes the permutation
        order = [i for idxs in groups.values() for i in idxs]
        b["order"] = order
        slot_of = {img_i: k for k, img_i in enumerate(order)}
        dev_images = [dev_images[i] for i in order]
        groups = {hw: [slot_of[i] for i in idxs] for hw, idxs in groups.items()}
        for (h, w), idxs in groups.items():
            scales = pyramid_scales((h, w, 3), cfg)
            # same-sized images of the call as one (n, h, w, 3) tensor (slots of a group are consecutive)
            stacked = dev_images[idxs[0]].unsqueeze(0) if len(idxs) == 1 else torch.stack([dev_images[i] for i in idxs])
            for li, s in enumerate(scales):
                _, _, hp, wp = level_geometry(h, w, s, cfg.max_resolution)
                # activations of the widest layer: 64 ch x 4 B per pixel, a few tensors live at once
                per_image = hp * wp * 64 * 4 * 3 * nf
                chunk = max(1, min(len(idxs), 32 // nf, int(self.max_batch_bytes // max(1, per_image))))
                for c0 in range(0, len(idxs), chunk):
                    sub = idxs[c0:c0 + chunk]
                    data, info = self._level_batch(stacked[c0:c0 + chunk], s, flips)
                    self.net.forward_body(data, fast=self.net.use_fast(s))
                    self.net.run_tail_batched(nf, info, b["dets"], b["offs"], image_base=sub[0], passes_total=passes,
                                              pass_base=li * nf, det_cap=b["cap"], det_thresh=cfg.thresh)
        method = 1 if cfg.nms_method == "BBOX_VOTE" else 0
        if cfg.nms_method not in ("BBOX_VOTE", "NMS"):
            raise NotImplementedError("Unknown NMS method: {}".format(cfg.nms_method))     # lib/test.py:173-175
        if self._post_stream is None:
            self._post_stream = torch.cuda.Stream(device=self.device)
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self._post_stream):
            self._post_stream.wait_event(ready)
            torch.add(b["seg_begin"], b["offs"][:, passes], out=b["seg_end"])
            L.call("shf_postprocess", _ptr(b["dets"]), _ptr(b["seg_begin"]), _ptr(b["seg_end"]), B, b["cap"],
                   float(cfg.nms_thresh), method, int(cfg.nms_mode), _ptr(b["out_idx"]), _ptr(b["out_dets"]),
                   _ptr(b["out_count"]), b["out_cap"], _ptr(b["ws"]), b["ws_bytes"], _stream())
            done = torch.cuda.Event()
            done.record()
        b["done"] = done
        self.net.launches += 7               # seg_end add + det_keys, sort, sorted boxes, IoU masks, sweep, reduce / emit
        return b

    def run_after_results(self, b, fn):
        """Runs ``fn()`` on the post-processing stream, ordered after the vote / NMS of ``b`` (e.g. the all-gather of the
        boxes in a multi-GPU run, so that it too overlaps the next batch); ``wait_results`` then also covers it."""
        with torch.cuda.stream(self._post_stream):
            out = fn()
            done = torch.cuda.Event()
            done.record()
        b["done"] = done
        return out

    def wait_results(self, b):
        """Orders the current stream after the post-processing of ``b`` (call before reading out_dets / out_idx /
        out_count on the device; ``download`` does it itself)."""
        if b.get("done") is not None:
            torch.cuda.current_stream().wait_event(b["done"])

    def download(self, b, B: int) -> List[np.ndarray]:
        """Device results -> list of (M,5) arrays ``[x1,y1,x2,y2,score]`` (float64 for BBOX_VOTE, as
        ``bbox_vote`` returns; float32 rows of ``dets`` for NMS)."""
        self.wait_results(b)
        counts = b["out_count"][:B].cpu().numpy()
        # the range guard rides on the synchronisation the counts just paid for; a fast-format level outside its
        # exponent window disables the format (sticky) and the call is repeated on split fp16
        if self.net.check_ranges(b["guard"]) and b.get("images") is not None and not b.get("retried"):
            b2 = self.detect_device(b["images"])
            b2["retried"] = True
            return self.download(b2, B)
        if int(counts.max(initial=0)) > b["out_cap"]:
            raise L.ShfError("post-processing produced %d rows for one image but the buffer holds %d"
                             % (int(counts.max()), b["out_cap"]))
        slots = []
        if self.cfg.nms_method == "BBOX_VOTE":
            host = b["out_dets"][:B, :max(1, int(counts.max(initial=0)))].cpu().numpy()
            for k in range(B):
                slots.append(host[k, :counts[k]].astype(np.float64))
        else:
            for k in range(B):
                idx = b["out_idx"][k, :int(counts[k])].long()
                slots.append(b["dets"][k].index_select(0, idx).cpu().numpy())
        out = [None] * B
        for k, img_i in enumerate(b.get("order", range(B))):
            out[img_i] = slots[k]
        return out

    def detect(self, images: List[np.ndarray]) -> List[np.ndarray]:
        dev = self.upload(images)
        b = self.detect_device(dev)
        return self.download(b, len(images))

    def raw_detections(self, b, i: int) -> np.ndarray:
        """Pre-vote detections of image i (concatenated passes, score > thresh), for parity checks."""
        k = list(b.get("order", range(i + 1))).index(i)
        n = int(b["offs"][k, -1].item())
        return b["dets"][k, :n].cpu().numpy()
