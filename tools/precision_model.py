"""CPU model of the tensor-core operand formats (runs without a GPU; uses oracle/ as the fp32 yardstick).

Question it answers: how far do the final boxes / scores of one pyramid level move, in raw-image px, when every
tcgen05 convolution consumes operands in a given reduced format?  Products are accumulated in float64 here, so
only the OPERAND rounding is modelled (the accumulation error of the tensor core is measured on the GPU).

schemes:
  exact      fp32 operands (yardstick against itself: 0)
  f16        single fp16 operand planes                       x ~ rn16(x), w ~ rn16(w)
  h2         split fp16 (hi + lo, 3 MMAs: hi*hi + hi*lo + lo*hi)
  h2f8v2     SHIPPED hf8: fp16 main term + fp8 correction term with fixed exponents:
               x = ah + al,  ah = rn16(x),  al8 = e4m3(al * 2^6),  ah8 = e4m3(ah * 2^-5)
               w*2^k = wh + wl, wh = rn16,  wh8 = e4m3(wh * 2^-6), wl8 = e4m3(wl * 2^5)
               y = ah*wh  +  al8*wh8 + ah8*wl8        (one kind::f16 MMA + one K=32 kind::f8f6f4 MMA per 16 channels)
  h2f8       the first version: e5m2 activation bytes (al * 2^10, ah), weights e4m3(wh * 2^-10), e4m3(wl)
  h2f8c / h2f8e4 / h2f8e4b / cal   other e4m3 windows, incl. per-layer calibrated ones (study only)
 schemes that would BREAK the 2-MMAs-per-MAC ceiling (study only; DESIGN.md "what would it take to pass 0.50"):
  h15        1.5 MMAs: fp16 main term + ONE f8 correction al8*wh8 over 32 channels per instruction; the weight residual
             term ah*wl is dropped, i.e. weights are plain fp16
  h15q       1.5 MMAs: fp16 main term + both correction terms in 4-bit e2m1 ([al4 | ah4] x [wh4 | wl4], kind::mxf4 K = 64)
  wino       Winograd F(2x2, 3x3) (2.25x fewer MACs) with the shipped hf8 operand split applied in the TRANSFORM domain:
             V = B^T d B and U = G g G^T are split like activations / weights, M = sum_c U.V, Y = A^T M A in fp32
             (3x3 stride-1 dilation-1 convs only; the dilated heads and the 1x1 convs stay direct hf8)

usage: python tools/precision_model.py [level ...]      (levels = TEST.SCALES entries, default 100 300)
"""
import math
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from smallhardface_b200 import deploy
from oracle.net import OracleNet
from oracle import detect as OD, preprocess as PRE, layers as OL

F64 = torch.float64


def rn16(t):
    return t.to(torch.float16).to(F64)


def e5m2(t):
    return t.to(torch.float32).to(torch.float8_e5m2).to(F64)


def e4m3(t):
    return t.to(torch.float32).clamp(-448, 448).to(torch.float8_e4m3fn).to(F64)


def e2m1(t):
    """4-bit float (1 sign, 2 exponent, 1 mantissa): magnitudes {0, .5, 1, 1.5, 2, 3, 4, 6}, round to nearest, saturating."""
    grid = torch.tensor([0.0, 0.5, 1.0, 1.5, 2.0, 3.0, 4.0, 6.0], dtype=F64)
    a = t.to(F64).abs().clamp(max=6.0)
    idx = (a.unsqueeze(-1) - grid).abs().argmin(dim=-1)
    return torch.sign(t.to(F64)) * grid[idx]


# Winograd F(2x2, 3x3) matrices (Lavin & Gray)
_BT = torch.tensor([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], dtype=F64)
_G = torch.tensor([[1, 0, 0], [0.5, 0.5, 0.5], [0.5, -0.5, 0.5], [0, 0, 1]], dtype=F64)
_AT = torch.tensor([[1, 1, 1, 0], [0, 1, -1, -1]], dtype=F64)


def winograd_hf8(xt, ws):
    """3x3 / pad 1 / stride 1 convolution of xt (1,C,H,W) with ws (K,C,3,3) (already scaled by 2^k) through
    F(2x2,3x3), operands split hf8-style in the transform domain, products accumulated in float64."""
    _, C, H, W = xt.shape
    K = ws.shape[0]
    Hp, Wp = H + (H % 2), W + (W % 2)
    xp = torch.nn.functional.pad(xt, (1, 1 + Wp - W, 1, 1 + Hp - H))
    tiles = xp.unfold(2, 4, 2).unfold(3, 4, 2)                       # (1, C, th, tw, 4, 4)
    th, tw = tiles.shape[2], tiles.shape[3]
    d = tiles.reshape(C, th * tw, 4, 4)
    V = _BT @ d @ _BT.t()                                            # (C, T, 4, 4)
    U = _G @ ws @ _G.t()                                             # (K, C, 4, 4)
    vh = rn16(V)
    val8, vah8 = e4m3((V - vh) * 64.0), e4m3(vh / 32.0)
    uh = rn16(U)
    uh8, ul8 = e4m3(uh / 64.0), e4m3((U - uh) * 32.0)
    M = (torch.einsum("kcij,ctij->ktij", uh, vh) + torch.einsum("kcij,ctij->ktij", uh8, val8)
         + torch.einsum("kcij,ctij->ktij", ul8, vah8))
    M = M.to(torch.float32).to(F64)                                  # accumulators are fp32
    Y = _AT @ M @ _AT.t()                                            # (K, T, 2, 2)
    y = Y.reshape(K, th, tw, 2, 2).permute(0, 1, 3, 2, 4).reshape(1, K, th * 2, tw * 2)
    return y[:, :, :H, :W]


def quant_act(x, scheme):
    """value the NEXT consumer sees for an activation stored in `scheme`'s format"""
    x = x.to(F64)
    if scheme == "exact":
        return x
    ah = rn16(x)
    if scheme == "f16":
        return ah
    al = x - ah
    if scheme == "h2":
        return ah + rn16(al)
    if scheme == "h2f8" or scheme.startswith("mix") or scheme.startswith("cal"):
        return ah + e5m2(al * 1024.0) / 1024.0      # (cal: only used for the non-tcgen05 consumers; approximation)
    if scheme in ("h2f8v2", "h15", "wino"):     # e4m3 bytes with fixed exponents (al * 2^6, hi * 2^-5), saturating
        return ah + e4m3(al * 64.0) / 64.0
    if scheme == "h15q":                         # e2m1 residual: al * 2^9 puts |al| <= 2^-11 |x| (|x| ~ 2^10) at ~1..6
        return ah + e2m1(al * 512.0) / 512.0
    if scheme in ("h2f8e4", "h2f8e4b"):          # e4m3 residual with a static 2^9 scale (full precision for 2^-4 <= |x| < 2^11)
        return ah + e4m3(al * 512.0) / 512.0
    if scheme == "h2f8c":           # e5m2 residual, e4m3 copy of hi
        return ah + e5m2(al * 1024.0) / 1024.0
    raise ValueError(scheme)


# ---- experimental (not shipped): e4m3 activation bytes with per-layer exponent windows from a calibration pass ----
CAL = {"max": {}, "mode": None, "headroom": 4.0}      # layer index -> max |x| seen by that conv's input


def e4m3_sat(t):
    return t.to(torch.float32).clamp(-448, 448).to(torch.float8_e4m3fn).to(F64)


def make_conv(scheme):
    real_conv = OL.conv
    mix_from = None
    if scheme.startswith("mix"):                 # "mixN": tcgen05 convs 0..N-1 on h2f8, N.. on h2 (dilated net: 15 = dim_red, 16.. = heads)
        mix_from = int(scheme[3:])
    counter = [0]

    def conv(x, w, b=None, pad=(0, 0), stride=(1, 1), dilation=(1, 1), group=1, engine="sgemm", **kw):
        cin = w.shape[1]
        nonlocal scheme
        if mix_from is not None and cin % 64 == 0 and w.shape[0] % 64 == 0:
            scheme = "h2f8" if counter[0] < mix_from else "h2"
            counter[0] += 1
        if scheme in ("cal", "calrec") and cin % 64 == 0 and w.shape[0] % 64 == 0:
            li = counter[0]
            counter[0] += 1
            xt = torch.from_numpy(np.ascontiguousarray(x)).to(F64)
            if scheme == "calrec":                     # calibration pass: record max |x| per layer, compute exactly
                CAL["max"][li] = max(CAL["max"].get(li, 0.0), float(xt.abs().max()))
                return real_conv(x, w, b, pad, stride, dilation, group, engine="torch")
            xmax = CAL["max"][li] * CAL["headroom"]
            # window: al <= 2^-11 * xmax -> al * 2^p <= 256 ; ah <= xmax -> ah * 2^-r <= 256
            p_exp = math.floor(math.log2(256.0 / (xmax * 2.0 ** -11)))
            r_exp = math.ceil(math.log2(xmax / 256.0))
            wt = torch.from_numpy(np.ascontiguousarray(w)).to(F64)
            k = int(14 - math.ceil(math.log2(float(wt.abs().max()))))
            ws = wt * 2.0 ** k
            cv = lambda a, bb: torch.nn.functional.conv2d(a, bb, None, stride=stride, padding=pad, dilation=dilation)
            # the stored activation itself is ah + al8 / 2^p (what the previous layer wrote)
            ah = rn16(xt)
            al8 = e4m3_sat((xt - ah) * 2.0 ** p_exp)
            ah8 = e4m3_sat(ah * 2.0 ** -r_exp)
            wh = rn16(ws)
            wl = ws - wh
            wh8 = e4m3_sat(wh * 2.0 ** -p_exp) if p_exp <= 14 else e4m3_sat(wh * 2.0 ** -p_exp)
            wl8 = e4m3_sat(wl * 2.0 ** r_exp)
            y = (cv(ah, wh) + cv(al8, wh8) + cv(ah8, wl8)) * 2.0 ** (-k)
            if b is not None:
                y = y + torch.from_numpy(np.asarray(b)).to(F64)[None, :, None, None]
            return y.to(torch.float32).numpy()
        if cin % 64 or w.shape[0] % 64 or scheme == "exact":          # conv1_1, cls/bbox 1x1: fp32 SIMT kernels
            xq = quant_act(torch.from_numpy(np.ascontiguousarray(x)), scheme if cin % 64 == 0 else "exact")
            return real_conv(xq.to(torch.float32).numpy(), w, b, pad, stride, dilation, group, engine="torch")
        xt = torch.from_numpy(np.ascontiguousarray(x)).to(F64)
        wt = torch.from_numpy(np.ascontiguousarray(w)).to(F64)
        amax = float(wt.abs().max())
        k = int(14 - math.ceil(math.log2(amax)))
        ws = wt * 2.0 ** k
        cv = lambda a, bb: torch.nn.functional.conv2d(a, bb, None, stride=stride, padding=pad, dilation=dilation)
        ah = rn16(xt)
        al = xt - ah
        wh = rn16(ws)
        wl = ws - wh
        if scheme == "f16":
            y = cv(ah, wh)
        elif scheme == "h2":
            all_, wll = rn16(al), rn16(wl)
            y = cv(ah, wh) + cv(ah, wll) + cv(all_, wh)
        elif scheme == "h2f8v2":
            y = cv(ah, wh) + cv(e4m3(al * 64.0), e4m3(wh / 64.0)) + cv(e4m3(ah / 32.0), e4m3(wl * 32.0))
        elif scheme == "h15":
            y = cv(ah, wh) + cv(e4m3(al * 64.0), e4m3(wh / 64.0))
        elif scheme == "h15q":
            # 4-bit planes: al * 2^9, ah * 2^-8 (|x| up to ~1500 -> ~6), wh * 2^-12, wl * 2^5 (|wl| <= 4 at the 2^14 scale -> <= 6 after /... )
            y = cv(ah, wh) + cv(e2m1(al * 512.0), e2m1(wh / 4096.0)) * 8.0 + cv(e2m1(ah / 256.0), e2m1(wl)) * 256.0
        elif scheme == "wino":
            if w.shape[2:] == (3, 3) and tuple(dilation) == (1, 1) and tuple(pad) == (1, 1) and tuple(stride) == (1, 1):
                y = winograd_hf8(quant_act(xt, "h2f8v2"), ws)
            else:
                y = cv(ah, wh) + cv(e4m3(al * 64.0), e4m3(wh / 64.0)) + cv(e4m3(ah / 32.0), e4m3(wl * 32.0))
        elif scheme in ("h2f8", "h2f8e4", "h2f8e4b", "h2f8c"):
            if scheme == "h2f8":
                al8 = e5m2(al * 1024.0)
                ah8 = e5m2(ah)
            elif scheme == "h2f8c":
                al8 = e5m2(al * 1024.0)
                ah8 = e4m3(ah / 256.0) * 256.0
            elif scheme == "h2f8e4":
                al8 = e4m3(al * 512.0) * 2.0
                ah8 = e5m2(ah)
            else:
                al8 = e4m3(al * 512.0) * 2.0
                ah8 = e4m3(ah / 256.0) * 256.0
            wh8 = e4m3(wh / 1024.0)
            wl8 = e4m3(wl)
            y = cv(ah, wh) + cv(al8, wh8) + cv(ah8, wl8)
        else:
            raise ValueError(scheme)
        y = y * 2.0 ** (-k)
        if b is not None:
            y = y + torch.from_numpy(np.asarray(b)).to(F64)[None, :, None, None]
        return y.to(torch.float32).numpy()

    return conv


def run_level(onet, im, lv, scheme):
    s = PRE.pyramid_scales(im.shape, (lv, lv + 1))[0]
    blob = PRE.get_image_blobs(im, [s])[0]
    saved = OL.conv
    OL.conv = make_conv(scheme)
    try:
        p, bx = OD.forward_level(onet, blob, s)
    finally:
        OL.conv = saved
    return s, p, bx, onet.last_order.copy()


def run_224(onet, scheme):
    im = np.random.RandomState(3).randint(0, 256, (224, 224, 3)).astype(np.uint8)
    data = np.ascontiguousarray((im.astype(np.float32) - np.array([[[102.9801, 115.9465, 122.7717]]])).astype(np.float32).transpose(2, 0, 1)[None])
    saved = OL.conv
    OL.conv = make_conv(scheme)
    try:
        p, bx = OD.forward_level(onet, data, 1.0)
    finally:
        OL.conv = saved
    return 1.0, p, bx, onet.last_order.copy()


def main():
    levels = [int(a) for a in sys.argv[1:]] or [100, 300]
    proto, model = deploy.write_synthetic_deployment(os.path.join(tempfile.gettempdir(), "shf_b200_deploy"), dilation=True)
    onet = OracleNet(proto, model, engine="torch", fast=True)
    im = deploy.synthetic_image(3)
    for lv in levels:
        runner = (lambda sch: run_224(onet, sch)) if lv == 224 else (lambda sch: run_level(onet, im, lv, sch))
        s, p0, b0, o0 = runner("exact")
        for scheme in (os.environ.get("SCHEMES", "f16,h2,h2f8,h2f8c,h2f8e4,h2f8e4b").split(",")):
            _, p1, b1, o1 = runner(scheme)
            # align rows by anchor index (order = anchor ids in descending score order)
            m0 = {int(a): i for i, a in enumerate(o0)}
            idx = [(m0[int(a)], j) for j, a in enumerate(o1) if int(a) in m0]
            i0 = np.array([a for a, _ in idx]); i1 = np.array([b for _, b in idx])
            db = np.abs(b0[i0] - b1[i1]).max() if len(idx) else float("nan")
            ds = np.abs(p0[i0] - p1[i1]).max() if len(idx) else float("nan")
            print("level %4d scale %.4f  %-7s rows %5d/%5d common %5d  score err %.2e  box err %.2e raw px (%.2e level px)"
                  % (lv, s, scheme, len(o1), len(o0), len(idx), ds, db, db * s), flush=True)


if __name__ == "__main__":
    main()
