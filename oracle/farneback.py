"""CPU restatement of OpenCV's dense optical flow (Farneback) as the FLF channel selector calls it.

TEST INFRASTRUCTURE ONLY (nothing under worldforge_b200/ imports this).  The reference scores the 16 latent channels with
``cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0)`` (reference
wan_for_worldforge/utils/scheduling_unipc_multistep_clean.py:220-224) - a third-party dependency (opencv-python, not
vendored in /root/reference; 4.13.0 in this image).  This file restates the published algorithm of
``modules/video/src/optflowgf.cpp`` (FarnebackPolyExp, FarnebackUpdateMatrices, FarnebackUpdateFlow_Blur and the
pyramid driver) in numpy so that a device kernel can be written and checked against something that is not a black box
(the GPU scoring is gated on selection-set equality with OpenCV).  It is pinned against OpenCV
itself in tests/test_farneback_oracle.py.

Facts of the call that matter (they fall out of the driver loop):
* latent frames are 60 x 104 (480p) or 90 x 160 (720p): the pyramid loop stops at the first level whose smaller side would
  drop under 32 pixels, so 60 x 104 runs ONE level (full resolution) and 90 x 160 two (45 x 80 then 90 x 160; sigma of the
  half-resolution level = (1/0.5 - 1)/2 = 0.5, kernel size cvRound(2.5)|1 = 3);
* at scale 1 the pre-smoothing is GaussianBlur(3 x 3, sigma 0) = the fixed [1/4, 1/2, 1/4] kernel, reflect-101 border;
* flags = 0: the flow update uses a 15 x 15 BOX window (replicated border), three iterations per level, matrices refreshed
  after the first two.
"""
from __future__ import annotations

import numpy as np

_BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], dtype=np.float32)


def _prepare_gaussian(n: int, sigma: float):
    """FarnebackPrepareGaussian: the 1-D weights g, x*g, x^2*g (float32) and the four entries of G^-1 that are used (double)."""
    if sigma < np.finfo(np.float32).eps:
        sigma = n * 0.3
    x = np.arange(-n, n + 1)
    g = np.exp(-x * x / (2 * sigma * sigma)).astype(np.float32)
    s = 1.0 / float(np.sum(g.astype(np.float64)))
    g = (g.astype(np.float64) * s).astype(np.float32)
    xg = (x * g.astype(np.float64)).astype(np.float32)
    xxg = (x * x * g.astype(np.float64)).astype(np.float32)
    G = np.zeros((6, 6))
    gd = g.astype(np.float64)
    for yy in range(-n, n + 1):
        for xx in range(-n, n + 1):
            w = gd[yy + n] * gd[xx + n]
            G[0, 0] += w
            G[1, 1] += w * xx * xx
            G[3, 3] += w * xx * xx * xx * xx
            G[5, 5] += w * xx * xx * yy * yy
    G[2, 2] = G[0, 3] = G[0, 4] = G[3, 0] = G[4, 0] = G[1, 1]
    G[4, 4] = G[3, 3]
    G[3, 4] = G[4, 3] = G[5, 5]
    inv = np.linalg.inv(G)
    return g, xg, xxg, inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]


def poly_exp(src: np.ndarray, n: int = 5, sigma: float = 1.2) -> np.ndarray:
    """FarnebackPolyExp: float32 [H, W] -> float32 [H, W, 5] = coefficients (y, x, yy, xx, xy) of the local quadratic."""
    H, W = src.shape
    g, xg, xxg, ig11, ig03, ig33, ig55 = _prepare_gaussian(n, sigma)
    src = src.astype(np.float32)
    # vertical pass (float32 accumulation, rows replicated at the border)
    r0 = src * g[n]
    r1 = np.zeros_like(src)
    r2 = np.zeros_like(src)
    rows = np.arange(H)
    for k in range(1, n + 1):
        up = src[np.maximum(rows - k, 0)]
        dn = src[np.minimum(rows + k, H - 1)]
        p = up + dn
        r0 = (r0 + g[n + k] * p).astype(np.float32)
        r1 = (r1 + xg[n + k] * (dn - up)).astype(np.float32)
        r2 = (r2 + xxg[n + k] * p).astype(np.float32)
    # horizontal pass (double accumulation, columns replicated)
    cols = np.arange(W)
    b1 = r0.astype(np.float64) * g[n]
    b3 = r1.astype(np.float64) * g[n]
    b5 = r2.astype(np.float64) * g[n]
    b2 = np.zeros((H, W))
    b4 = np.zeros((H, W))
    b6 = np.zeros((H, W))
    for k in range(1, n + 1):
        rp, rm = np.minimum(cols + k, W - 1), np.maximum(cols - k, 0)
        tg = (r0[:, rp] + r0[:, rm]).astype(np.float32).astype(np.float64)
        b1 += tg * g[n + k]
        b4 += tg * xxg[n + k]
        b2 += (r0[:, rp] - r0[:, rm]).astype(np.float32).astype(np.float64) * xg[n + k]
        b3 += (r1[:, rp] + r1[:, rm]).astype(np.float32).astype(np.float64) * g[n + k]
        b6 += (r1[:, rp] - r1[:, rm]).astype(np.float32).astype(np.float64) * xg[n + k]
        b5 += (r2[:, rp] + r2[:, rm]).astype(np.float32).astype(np.float64) * g[n + k]
    out = np.empty((H, W, 5), dtype=np.float32)
    out[..., 1] = b2 * ig11
    out[..., 0] = b3 * ig11
    out[..., 3] = b1 * ig03 + b4 * ig33
    out[..., 2] = b1 * ig03 + b5 * ig33
    out[..., 4] = b6 * ig55
    return out


def update_matrices(R0: np.ndarray, R1: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """FarnebackUpdateMatrices over the whole image: the per-pixel 2x2 system (g11, g12, g22, h1, h2), float32."""
    H, W, _ = R0.shape
    f32 = np.float32
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    dx, dy = flow[..., 0].astype(f32), flow[..., 1].astype(f32)
    fx = (xs.astype(f32) + dx).astype(f32)
    fy = (ys.astype(f32) + dy).astype(f32)
    x1 = np.floor(fx).astype(np.int64)
    y1 = np.floor(fy).astype(np.int64)
    fx = (fx - x1.astype(f32)).astype(f32)
    fy = (fy - y1.astype(f32)).astype(f32)
    inside = (x1 >= 0) & (x1 < W - 1) & (y1 >= 0) & (y1 < H - 1)
    xc, yc = np.clip(x1, 0, W - 2), np.clip(y1, 0, H - 2)
    one = f32(1.0)
    a00 = ((one - fx) * (one - fy)).astype(f32)
    a01 = (fx * (one - fy)).astype(f32)
    a10 = ((one - fx) * fy).astype(f32)
    a11 = (fx * fy).astype(f32)

    def bil(c):
        v = (a00 * R1[yc, xc, c]).astype(f32)
        v = (v + a01 * R1[yc, xc + 1, c]).astype(f32)
        v = (v + a10 * R1[yc + 1, xc, c]).astype(f32)
        return (v + a11 * R1[yc + 1, xc + 1, c]).astype(f32)

    r2 = np.where(inside, bil(0), f32(0))
    r3 = np.where(inside, bil(1), f32(0))
    r4 = np.where(inside, ((R0[..., 2] + bil(2)) * f32(0.5)).astype(f32), R0[..., 2])
    r5 = np.where(inside, ((R0[..., 3] + bil(3)) * f32(0.5)).astype(f32), R0[..., 3])
    r6 = np.where(inside, ((R0[..., 4] + bil(4)) * f32(0.25)).astype(f32), (R0[..., 4] * f32(0.5)).astype(f32))
    r2 = ((R0[..., 0] - r2) * f32(0.5)).astype(f32)
    r3 = ((R0[..., 1] - r3) * f32(0.5)).astype(f32)
    r2 = (r2 + ((r4 * dy).astype(f32) + (r6 * dx).astype(f32)).astype(f32)).astype(f32)
    r3 = (r3 + ((r6 * dy).astype(f32) + (r5 * dx).astype(f32)).astype(f32)).astype(f32)
    B = len(_BORDER)

    # (x < B ? border[x] : 1) * (x >= W-B ? border[W-x-1] : 1) * (y < B ? ...) * (y >= H-B ? ...), evaluated left to right
    sx_lo = np.where(xs < B, _BORDER[np.minimum(xs, B - 1)], one)
    sx_hi = np.where(xs >= W - B, _BORDER[np.clip(W - xs - 1, 0, B - 1)], one)
    sy_lo = np.where(ys < B, _BORDER[np.minimum(ys, B - 1)], one)
    sy_hi = np.where(ys >= H - B, _BORDER[np.clip(H - ys - 1, 0, B - 1)], one)
    scale = (((sx_lo * sx_hi).astype(f32) * sy_lo).astype(f32) * sy_hi).astype(f32)
    r2, r3, r4, r5, r6 = [(v * scale).astype(f32) for v in (r2, r3, r4, r5, r6)]
    M = np.empty((H, W, 5), dtype=f32)
    M[..., 0] = ((r4 * r4).astype(f32) + (r6 * r6).astype(f32)).astype(f32)
    M[..., 1] = ((r4 + r5).astype(f32) * r6).astype(f32)
    M[..., 2] = ((r5 * r5).astype(f32) + (r6 * r6).astype(f32)).astype(f32)
    M[..., 3] = ((r4 * r2).astype(f32) + (r6 * r3).astype(f32)).astype(f32)
    M[..., 4] = ((r6 * r2).astype(f32) + (r5 * r3).astype(f32)).astype(f32)
    return M


def box_flow(M: np.ndarray, block_size: int) -> np.ndarray:
    """The solve of FarnebackUpdateFlow_Blur: 15 x 15 box sums of M with replicated borders (double), then the 2x2 solve."""
    H, W, _ = M.shape
    m = block_size // 2
    Md = M.astype(np.float64)
    rows = np.arange(H)
    v = np.zeros_like(Md)
    for d in range(-m, m + 1):
        v += Md[np.clip(rows + d, 0, H - 1)]
    cols = np.arange(W)
    s = np.zeros_like(Md)
    for d in range(-m, m + 1):
        s += v[:, np.clip(cols + d, 0, W - 1)]
    s *= 1.0 / (block_size * block_size)
    g11, g12, g22, h1, h2 = (s[..., i] for i in range(5))
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    flow = np.empty((H, W, 2), dtype=np.float32)
    flow[..., 0] = (g11 * h2 - g12 * h1) * idet
    flow[..., 1] = (g22 * h1 - g12 * h2) * idet
    return flow


def _gauss3_reflect101(img: np.ndarray) -> np.ndarray:
    """GaussianBlur(3x3, sigma 0) on float32: the fixed [1/4, 1/2, 1/4] kernel, BORDER_REFLECT_101, rows then columns."""
    f32 = np.float32
    p = np.pad(img.astype(f32), ((0, 0), (1, 1)), mode="reflect")
    r = (p[:, :-2] * f32(0.25) + p[:, 1:-1] * f32(0.5)).astype(f32)
    r = (r + p[:, 2:] * f32(0.25)).astype(f32)
    p = np.pad(r, ((1, 1), (0, 0)), mode="reflect")
    c = (p[:-2] * f32(0.25) + p[1:-1] * f32(0.5)).astype(f32)
    return (c + p[2:] * f32(0.25)).astype(f32)


def _blur3(img: np.ndarray, k) -> np.ndarray:
    """Separable 3-tap blur, BORDER_REFLECT_101, rows then columns, float32 (GaussianBlur with a 3 x 3 kernel)."""
    f32 = np.float32
    p = np.pad(img.astype(f32), ((0, 0), (1, 1)), mode="reflect")
    r = (p[:, :-2] * f32(k[0]) + p[:, 1:-1] * f32(k[1])).astype(f32)
    r = (r + p[:, 2:] * f32(k[2])).astype(f32)
    p = np.pad(r, ((1, 1), (0, 0)), mode="reflect")
    c = (p[:-2] * f32(k[0]) + p[1:-1] * f32(k[1])).astype(f32)
    return (c + p[2:] * f32(k[2])).astype(f32)


def _gauss_kernel3(sigma: float):
    """cv::getGaussianKernel(3, sigma > 0, CV_32F)."""
    e = np.exp(-1.0 / (2.0 * sigma * sigma))
    s = 1.0 / (1.0 + 2.0 * e)
    return np.array([e * s, s, e * s], dtype=np.float32)


def _half(img: np.ndarray) -> np.ndarray:
    """cv::resize(INTER_LINEAR) to exactly half the size (even sides): the mean of each 2 x 2 cell."""
    return ((img[0::2, 0::2] + img[0::2, 1::2] + img[1::2, 0::2] + img[1::2, 1::2]) * np.float32(0.25)).astype(np.float32)


def _double(x: np.ndarray) -> np.ndarray:
    """cv::resize(INTER_LINEAR) to exactly twice the size: source coordinate (dst + 0.5) / 2 - 0.5, clamped taps."""
    H, W = x.shape[:2]

    def axis(n):
        src = (np.arange(2 * n) + 0.5) / 2 - 0.5
        i0 = np.floor(src).astype(np.int64)
        w = (src - i0).astype(np.float32)
        return np.clip(i0, 0, n - 1), np.clip(i0 + 1, 0, n - 1), w

    y0, y1, wy = axis(H)
    x0, x1, wx = axis(W)
    wx, wy = wx[None, :, None], wy[:, None, None]
    top = x[y0][:, x0] * (1 - wx) + x[y0][:, x1] * wx
    bot = x[y1][:, x0] * (1 - wx) + x[y1][:, x1] * wx
    return (top * (1 - wy) + bot * wy).astype(np.float32)


def farneback(prev: np.ndarray, nxt: np.ndarray, pyr_scale: float = 0.5, levels: int = 3, winsize: int = 15,
              iterations: int = 3, poly_n: int = 5, poly_sigma: float = 1.2) -> np.ndarray:
    """uint8 [H, W] x 2 -> float32 flow [H, W, 2] (dx, dy); flags = 0.  Pyramids of one level (min side * 0.5 < 32: the
    60 x 104 latent frames of 480p - reproduced to a few 1e-6 px) and of two levels with even sides (90 x 160, 720p: the
    half-resolution level is GaussianBlur(3 x 3, sigma 0.5) + a 2 x 2 mean, its flow comes back through a bilinear doubling
    times 2 - reproduced to ~1e-4 px because OpenCV's filter engine rounds the blur differently).  Deeper pyramids raise."""
    H, W = prev.shape
    scale, k = 1.0, 0
    while k < levels:
        scale *= pyr_scale
        if W * scale < 32 or H * scale < 32:
            break
        k += 1
    if k > 1 or (k == 1 and (pyr_scale != 0.5 or H % 2 or W % 2)):
        raise NotImplementedError(f"{H}x{W}: the pyramid has {k + 1} levels; one level, or two with even sides at scale 0.5, are restated")
    flow = None
    for lvl in range(k, -1, -1):
        if lvl == 0:
            imgs = [_gauss3_reflect101(img.astype(np.float32)) for img in (prev, nxt)]
        else:
            imgs = [_half(_blur3(img.astype(np.float32), _gauss_kernel3(0.5))) for img in (prev, nxt)]
        R = [poly_exp(im, poly_n, poly_sigma) for im in imgs]
        h, w = imgs[0].shape
        flow = np.zeros((h, w, 2), dtype=np.float32) if flow is None else (_double(flow) * np.float32(1.0 / pyr_scale)).astype(np.float32)
        M = update_matrices(R[0], R[1], flow)
        for i in range(iterations):
            flow = box_flow(M, winsize)
            if i < iterations - 1:
                M = update_matrices(R[0], R[1], flow)
    return flow
