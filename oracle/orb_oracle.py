"""CPU restatement of the reference's ORB feature extraction -- TEST INFRASTRUCTURE ONLY (same rules as oracle/__init__.py).

What it restates: the ``detector == 3`` branch of ``StructFromMotion::getFeature`` (/root/reference/src/Sfm.cpp:358-384)

    cv::ORB::create(nfeatures=500, scaleFactor=1.2f, nlevels=8, edgeThreshold=31, firstLevel=0, WTA_K=2,
                    scoreType=HARRIS_SCORE, patchSize=31, fastThreshold=20)         src/Sfm.cpp:360-368
    detector->detectAndCompute(image, noArray(), kps, descriptors, false)            src/Sfm.cpp:373

whose arithmetic lives in OpenCV (features2d/src/orb.cpp, fast.cpp, fast_score.cpp; imgproc resize / GaussianBlur), not vendored
under /root/reference.  The published algorithm (Rublee et al., "ORB: an efficient alternative to SIFT or SURF", ICCV 2011, as
implemented by OpenCV), restated here in numpy with integer / fp32 arithmetic spelled out so that it can be reproduced bit for bit:

  1. pyramid: level l has size round(W / 1.2^l) x round(H / 1.2^l); level l is resized from level l-1 with INTER_LINEAR_EXACT
     (8.8 fixed-point bilinear, coefficients from double arithmetic); every level is extended by reflect-101 borders;
  2. per level: FAST-9/16 corners with threshold 20 and 3x3 non-maximum suppression on the corner score, corners closer than 31
     pixels to the level's border dropped, the 2N strongest (by FAST score, ties kept) retained, N = the level's share of 500;
  3. Harris response (7x7 block, k = 0.04) at the survivors, the N strongest (ties kept) retained;
  4. orientation by intensity centroid over a radius-15 disc, angle = fastAtan2(m01, m10) (OpenCV's degree-7 polynomial, fp32);
  5. level blurred with a 7x7 sigma-2 Gaussian: separable fp32 filter with fused multiply-adds, rounded to 8 bits at the end
     (NOT the 8.8 fixed-point path a stand-alone cv::GaussianBlur takes: ORB blurs a sub-matrix in place);
  6. rBRIEF: 256 intensity comparisons at the learned pattern's points rotated by the orientation (fp32, round-half-even).

The learned 256 x 4 pattern ("bit_pattern_31_") is a constant table of the published method; tools/extract_orb_pattern.py reads it
out of the cv2 wheel's binary into sfm_danpipeline_b200/orb_pattern.npy, which the product and this oracle both load.

Pin: tests/golden/temple_orb_features.npz (cv2 4.13.0 ORB on /root/reference/data/temple, made by tests/golden/make_golden.py);
tests/test_orb_oracle.py checks every stage of this file against cv2 when it is importable and the end result against the fixture.
Keypoint ORDER inside a level is not part of the contract: OpenCV's retainBest leaves it to std::nth_element.  The contract is
the keypoint SET per level and, per keypoint, angle / response / descriptor.
"""
from __future__ import annotations

import os

import numpy as np

N_FEATURES, SCALE_FACTOR, N_LEVELS, EDGE_THRESHOLD, PATCH_SIZE, FAST_THRESHOLD = 500, 1.2, 8, 31, 31, 20
HALF_PATCH = PATCH_SIZE // 2
HARRIS_BLOCK, HARRIS_K = 7, np.float32(0.04)
BORDER = 32  # max(edgeThreshold, ceil(halfPatch * sqrt 2), HARRIS_BLOCK_SIZE / 2) + 1

_HERE = os.path.dirname(os.path.abspath(__file__))
PATTERN_PATH = os.path.join(os.path.dirname(_HERE), "sfm_danpipeline_b200", "orb_pattern.npy")

# the 16-pixel Bresenham circle of radius 3, in OpenCV's order (dx, dy)
RING = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3), (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]



def gauss_kernel_7_s2() -> np.ndarray:
    """cv::getGaussianKernel(7, 2, CV_32F): exp(-(i-3)^2 / (2 sigma^2)) in double, normalised to sum 1, rounded to fp32
    (0x1.1f5f62p-4, 0x1.0c70fcp-3, 0x1.869472p-3, 0x1.ba95c0p-3, mirrored)."""
    x = np.arange(7, dtype=np.float64) - 3.0
    k = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    return (k / k.sum()).astype(np.float32)


GAUSS_7_S2 = gauss_kernel_7_s2()


def pattern() -> np.ndarray:
    return np.load(PATTERN_PATH)  # (256, 4) int32: x0, y0, x1, y1


# ----------------------------------------------------------------------------------------------------------------- 0. gray
def bgr_to_gray(bgr: np.ndarray) -> np.ndarray:
    """cv::cvtColor(BGR2GRAY) on 8-bit data: 15-bit fixed point, (B*3735 + G*19235 + R*9798 + 16384) >> 15 (probed against cv2)."""
    b, g, r = (bgr[..., i].astype(np.uint32) for i in range(3))
    return ((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------- 1. pyramid
def level_scales() -> np.ndarray:
    return np.array([np.float32(np.float64(np.float32(SCALE_FACTOR)) ** l) for l in range(N_LEVELS)], np.float32)


def level_sizes(w: int, h: int):
    return [(int(np.rint(w / s)), int(np.rint(h / s))) for s in level_scales()]  # cvRound(cols / scale): float division


def _linear_coeffs(src: int, dst: int):
    """INTER_LINEAR_EXACT taps of one axis: offset[d], c0[d], c1[d] in 8.8 fixed point."""
    scale = np.float64(1.0) / (np.float64(dst) / np.float64(src))
    ofs = np.zeros(dst, np.int64)
    c1 = np.zeros(dst, np.int64)
    for d in range(dst):
        f = scale * (np.float64(d) + 0.5) - 0.5
        i = int(np.floor(f))
        if i >= 0 and src > 1:
            if i < src - 1:
                ofs[d] = i
                c1[d] = int(np.rint((f - i) * 256.0))
            else:
                ofs[d], c1[d] = src - 1, 0  # past the last sample: the last sample
        else:
            ofs[d], c1[d] = 0, 0  # before the first sample: the first sample
    return ofs, 256 - c1, c1


def resize_linear_exact(img: np.ndarray, w: int, h: int) -> np.ndarray:
    sh, sw = img.shape
    ox, cx0, cx1 = _linear_coeffs(sw, w)
    oy, cy0, cy1 = _linear_coeffs(sh, h)
    a = img.astype(np.int64)
    ox1 = np.minimum(ox + 1, sw - 1)
    hor = a[:, ox] * cx0[None, :] + a[:, ox1] * cx1[None, :]          # 8.8
    oy1 = np.minimum(oy + 1, sh - 1)
    ver = hor[oy, :] * cy0[:, None] + hor[oy1, :] * cy1[:, None]      # 16.16
    return ((ver + (1 << 15)) >> 16).astype(np.uint8)


def pyramid(img: np.ndarray):
    h, w = img.shape
    levels = [img]
    for (lw, lh) in level_sizes(w, h)[1:]:
        levels.append(resize_linear_exact(levels[-1], lw, lh))
    return levels


def reflect101(img: np.ndarray, b: int = BORDER) -> np.ndarray:
    return np.pad(img, b, mode="reflect")


# ----------------------------------------------------------------------------------------------------------------- 2. FAST
def fast_scores(img: np.ndarray, threshold: int = FAST_THRESHOLD) -> np.ndarray:
    """Corner score of every pixel (0 where it is not a FAST-9 corner): the largest t for which the pixel still is a corner.
    score = max(max over the 16 arcs of min(v - p), max over arcs of min(p - v)) - 1, kept when > threshold - 1."""
    h, w = img.shape
    a = img.astype(np.int32)
    v = a[3:h - 3, 3:w - 3]
    d = np.stack([v - a[3 + dy:h - 3 + dy, 3 + dx:w - 3 + dx] for dx, dy in RING], 0)  # (16, h-6, w-6): centre minus ring
    d2 = np.concatenate([d, d[:8]], 0)
    best_dark = np.full(v.shape, -512, np.int32)    # arcs of ring pixels DARKER than the centre: min of d over the arc
    best_bright = np.full(v.shape, -512, np.int32)  # arcs BRIGHTER than the centre: min of -d
    for s in range(16):
        arc = d2[s:s + 9]
        best_dark = np.maximum(best_dark, arc.min(0))
        best_bright = np.maximum(best_bright, (-arc).min(0))
    m = np.maximum(best_dark, best_bright)
    score = np.zeros((h, w), np.int32)
    score[3:h - 3, 3:w - 3] = np.where(m > threshold, m - 1, 0)
    return score


def fast_keypoints(img: np.ndarray, threshold: int = FAST_THRESHOLD):
    """(x, y, score) of the corners that survive 3x3 non-maximum suppression (strictly greater than all 8 neighbours), row-major."""
    s = fast_scores(img, threshold)
    h, w = s.shape
    p = np.pad(s, 1)
    c = p[1:-1, 1:-1]
    keep = c > 0
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dx or dy:
                keep &= c > p[1 + dy:h + 1 + dy, 1 + dx:w + 1 + dx]
    ys, xs = np.nonzero(keep)
    return xs.astype(np.int32), ys.astype(np.int32), s[ys, xs].astype(np.float32)


def features_per_level():
    factor = np.float32(1.0) / np.float32(SCALE_FACTOR)
    n = np.float32(N_FEATURES) * (np.float32(1) - factor) / (np.float32(1) - np.float32(np.float64(factor) ** N_LEVELS))
    out, total = [], 0
    for _ in range(N_LEVELS - 1):
        k = int(np.rint(n))
        out.append(k)
        total += k
        n = np.float32(n * factor)
    out.append(max(N_FEATURES - total, 0))
    return out


def retain_best(resp: np.ndarray, n: int) -> np.ndarray:
    """KeyPointsFilter::retainBest as a SET: everything at least as strong as the n-th strongest (ties kept)."""
    if n >= len(resp):
        return np.ones(len(resp), bool)
    if n == 0:
        return np.zeros(len(resp), bool)
    kth = np.sort(resp)[::-1][n - 1]
    return resp >= kth


# ----------------------------------------------------------------------------------------------------------------- 3. Harris
def harris_responses(img: np.ndarray, xs, ys) -> np.ndarray:
    a = img.astype(np.int32)
    r = HARRIS_BLOCK // 2
    scale = np.float32(1.0) / (np.float32(4 * HARRIS_BLOCK) * np.float32(255.0))
    scale_sq_sq = np.float32(scale * scale * scale * scale)
    out = np.zeros(len(xs), np.float32)
    for i, (x0, y0) in enumerate(zip(xs, ys)):
        win = a[y0 - r - 1:y0 + r + 2, x0 - r - 1:x0 + r + 2]  # 9 x 9
        ix = (win[1:-1, 2:] - win[1:-1, :-2]) * 2 + (win[:-2, 2:] - win[:-2, :-2]) + (win[2:, 2:] - win[2:, :-2])
        iy = (win[2:, 1:-1] - win[:-2, 1:-1]) * 2 + (win[2:, :-2] - win[:-2, :-2]) + (win[2:, 2:] - win[:-2, 2:])
        A, B, C = np.float32(int((ix * ix).sum())), np.float32(int((iy * iy).sum())), np.float32(int((ix * iy).sum()))
        out[i] = np.float32(np.float32(np.float32(A * B) - np.float32(C * C)) - np.float32(np.float32(HARRIS_K * np.float32(A + B)) * np.float32(A + B))) * scale_sq_sq
    return out


# ----------------------------------------------------------------------------------------------------------------- 4. angle
def umax_table():
    vmax = int(np.floor(HALF_PATCH * np.sqrt(np.float32(2.0)) / 2 + 1))
    vmin = int(np.ceil(HALF_PATCH * np.sqrt(np.float32(2.0)) / 2))
    umax = np.zeros(HALF_PATCH + 2, np.int32)
    for v in range(vmax + 1):
        umax[v] = int(np.rint(np.sqrt(float(HALF_PATCH * HALF_PATCH - v * v))))
    v0 = 0
    for v in range(HALF_PATCH, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax


_P1 = np.float32(np.float32(0.9997878412794807) * np.float32(180 / np.pi))
_P3 = np.float32(np.float32(-0.3258083974640975) * np.float32(180 / np.pi))
_P5 = np.float32(np.float32(0.1555786518463281) * np.float32(180 / np.pi))
_P7 = np.float32(np.float32(-0.04432655554792128) * np.float32(180 / np.pi))
_EPS = np.float32(2.220446049250313e-16)


def fast_atan2(y, x) -> np.float32:
    y, x = np.float32(y), np.float32(x)
    ax, ay = np.float32(abs(x)), np.float32(abs(y))
    if ax >= ay:
        c = np.float32(ay / np.float32(ax + _EPS))
        c2 = np.float32(c * c)
        a = np.float32(np.float32(np.float32(np.float32(np.float32(np.float32(_P7 * c2) + _P5) * c2) + _P3) * c2 + _P1) * c)
    else:
        c = np.float32(ax / np.float32(ay + _EPS))
        c2 = np.float32(c * c)
        a = np.float32(np.float32(90.0) - np.float32(np.float32(np.float32(np.float32(np.float32(np.float32(_P7 * c2) + _P5) * c2) + _P3) * c2 + _P1) * c))
    if x < 0:
        a = np.float32(np.float32(180.0) - a)
    if y < 0:
        a = np.float32(np.float32(360.0) - a)
    return a


def ic_angles(img: np.ndarray, xs, ys) -> np.ndarray:
    a = img.astype(np.int64)
    umax = umax_table()
    out = np.zeros(len(xs), np.float32)
    us = np.arange(-HALF_PATCH, HALF_PATCH + 1)
    for i, (x, y) in enumerate(zip(xs, ys)):
        m01 = 0
        m10 = int((us * a[y, x - HALF_PATCH:x + HALF_PATCH + 1]).sum())
        for v in range(1, HALF_PATCH + 1):
            d = int(umax[v])
            plus, minus = a[y + v, x - d:x + d + 1], a[y - v, x - d:x + d + 1]
            u = np.arange(-d, d + 1)
            m10 += int((u * (plus + minus)).sum())
            m01 += v * int((plus - minus).sum())
        out[i] = fast_atan2(m01, m10)
    return out


# ----------------------------------------------------------------------------------------------------------------- 5. blur
def _fma32(a: np.ndarray, b, c: np.ndarray) -> np.ndarray:
    """fp32 fused multiply-add: the product of two fp32 numbers is exact in double, one rounding to double then to fp32
    (double rounding cannot bite: product 48 bits + addend 24 bits fit the 53-bit significand unless exponents are ~29 apart,
    where the addend is below half an ulp of the fp32 result either way)."""
    return (a.astype(np.float64) * np.float64(b) + c.astype(np.float64)).astype(np.float32)


def gaussian_blur_7x7(img: np.ndarray) -> np.ndarray:
    """cv::GaussianBlur(Size(7,7), 2, 2, BORDER_REFLECT_101) as ORB calls it: in place on a sub-matrix of its pyramid buffer,
    which rules out OpenCV's fixed-point 8-bit path and leaves the separable fp32 filter (probed against cv2.sepFilter2D, zero
    differing pixels over all 80 level images of data/temple):
      rows   : acc = k0*s[0]; acc = fma(s[i], k[i], acc) for i = 1..6                       (generic row filter, left to right)
      columns: acc = k3*r[3]; acc = fma(r[3-j] + r[3+j], k[3+j], acc) for j = 1..3          (symmetric column filter)
      out    = saturate_u8(round_half_even(acc))."""
    k = GAUSS_7_S2
    p = np.pad(img, 3, mode="reflect").astype(np.float32)
    h, w = img.shape
    hor = (k[0] * p[:, 0:w]).astype(np.float32)
    for i in range(1, 7):
        hor = _fma32(p[:, i:i + w], k[i], hor)
    ver = (k[3] * hor[3:3 + h]).astype(np.float32)
    for j in range(1, 4):
        ver = _fma32((hor[3 - j:3 - j + h] + hor[3 + j:3 + j + h]).astype(np.float32), k[3 + j], ver)
    return np.clip(np.rint(ver), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------------------------------------------- 6. rBRIEF
def brief_descriptors(blurred_ext: np.ndarray, xs, ys, angles_deg, border: int = BORDER) -> np.ndarray:
    """blurred_ext: the blurred level inside its reflect-101 frame of `border` pixels (unblurred frame, as in OpenCV's buffer)."""
    pat = pattern().astype(np.float32)
    out = np.zeros((len(xs), 32), np.uint8)
    for i, (x, y, ang) in enumerate(zip(xs, ys, angles_deg)):
        rad = np.float32(np.float32(ang) * np.float32(np.pi / 180.0))
        ca, sb = np.float32(np.cos(np.float64(rad))), np.float32(np.sin(np.float64(rad)))
        px0 = np.rint(pat[:, 0] * ca - pat[:, 1] * sb).astype(np.int32)
        py0 = np.rint(pat[:, 0] * sb + pat[:, 1] * ca).astype(np.int32)
        px1 = np.rint(pat[:, 2] * ca - pat[:, 3] * sb).astype(np.int32)
        py1 = np.rint(pat[:, 2] * sb + pat[:, 3] * ca).astype(np.int32)
        t0 = blurred_ext[y + border + py0, x + border + px0]
        t1 = blurred_ext[y + border + py1, x + border + px1]
        out[i] = np.packbits((t0 < t1).reshape(32, 8)[:, ::-1], axis=1).ravel()  # bit k of a byte = comparison k of its group of 8
    return out


# ----------------------------------------------------------------------------------------------------------------- all of it
def detect_and_compute(img: np.ndarray):
    """Returns (keypoints, descriptors): keypoints is a structured array (x, y, size, angle, response, octave) in level order,
    row-major inside a level; descriptors (n, 32) uint8 -- the rows imagesDescriptors[i] would hold (src/Sfm.cpp:381)."""
    levels = pyramid(img)
    scales = level_scales()
    quota = features_per_level()
    kp_dtype = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])
    all_kp, all_desc = [], []
    for l, lev in enumerate(levels):
        h, w = lev.shape
        xs, ys, resp = fast_keypoints(lev)
        inside = (xs >= EDGE_THRESHOLD) & (xs < w - EDGE_THRESHOLD) & (ys >= EDGE_THRESHOLD) & (ys < h - EDGE_THRESHOLD)
        xs, ys, resp = xs[inside], ys[inside], resp[inside]
        k = retain_best(resp, 2 * quota[l])
        xs, ys = xs[k], ys[k]
        hr = harris_responses(lev, xs, ys)
        k = retain_best(hr, quota[l])
        xs, ys, hr = xs[k], ys[k], hr[k]
        ext = reflect101(lev)
        ang = ic_angles(ext, xs + BORDER, ys + BORDER)
        blurred = reflect101(lev)
        blurred[BORDER:-BORDER, BORDER:-BORDER] = gaussian_blur_7x7(lev)
        desc = brief_descriptors(blurred, xs, ys, ang)
        kp = np.zeros(len(xs), kp_dtype)
        kp["x"], kp["y"] = xs.astype(np.float32) * scales[l], ys.astype(np.float32) * scales[l]
        kp["size"], kp["angle"], kp["response"], kp["octave"] = np.float32(PATCH_SIZE) * scales[l], ang, hr, l
        all_kp.append(kp)
        all_desc.append(desc)
    return np.concatenate(all_kp), np.concatenate(all_desc)
