"""Reference-faithful ORB extractor on the REAL OpenCV primitives (cv2).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows /root/reference/SingleRobotScenario/src/ORBextractor.cc line by line and
calls the same OpenCV functions the reference calls -- cv::resize, cv::copyMakeBorder,
cv::FAST (one call per 30-px cell, with the minThFAST retry), cv::GaussianBlur,
cv::fastAtan2 -- through cv2.  The non-OpenCV arithmetic (quad-tree, IC_Angle moments,
steered BRIEF) comes from orb_oracle.c.  This is what pins orb_oracle.c's restated
primitives, and it is the CPU baseline timed by bench.py (kind "port": the reference
C++ itself cannot be compiled here).
"""
import math
import numpy as np
import cv2

from . import (orb_params, level_size, distribute_octree, ic_moments, orb_descriptor, _empty_result)

EDGE_THRESHOLD = 19
PATCH_SIZE = 31


def compute_pyramid(P, image, with_border=False):
    """ORBextractor::ComputePyramid, ORBextractor.cc:1107-1132."""
    h, w = image.shape
    levels, bordered = [], []
    for level in range(P.nlevels):
        lw, lh = level_size(P, w, h, level)
        if level != 0:
            img = cv2.resize(levels[level - 1], (lw, lh), interpolation=cv2.INTER_LINEAR)
        else:
            img = image
        levels.append(img)
        if with_border:
            bordered.append(cv2.copyMakeBorder(img, EDGE_THRESHOLD, EDGE_THRESHOLD, EDGE_THRESHOLD,
                                               EDGE_THRESHOLD, cv2.BORDER_REFLECT_101))
    return (levels, bordered) if with_border else levels


def detect_cells_cv2(img, ini_th, min_th):
    """Cell loop of ComputeKeyPointsOctTree, ORBextractor.cc:769-829."""
    h, w = img.shape
    minBX = minBY = EDGE_THRESHOLD - 3
    maxBX, maxBY = w - EDGE_THRESHOLD + 3, h - EDGE_THRESHOLD + 3
    width, height = np.float32(maxBX - minBX), np.float32(maxBY - minBY)
    W = np.float32(30)
    nCols, nRows = int(width / W), int(height / W)
    out = []
    if nCols < 1 or nRows < 1:
        return np.zeros((0, 3), np.float32)
    wCell, hCell = int(math.ceil(width / np.float32(nCols))), int(math.ceil(height / np.float32(nRows)))
    det_ini = cv2.FastFeatureDetector_create(int(ini_th), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    det_min = cv2.FastFeatureDetector_create(int(min_th), True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for i in range(nRows):
        iniY = minBY + i * hCell
        maxY = iniY + hCell + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nCols):
            iniX = minBX + j * wCell
            maxX = iniX + wCell + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            cell = img[iniY:maxY, iniX:maxX]
            kps = det_ini.detect(cell)
            if len(kps) == 0:
                kps = det_min.detect(cell)
            for k in kps:
                out.append((k.pt[0] + j * wCell, k.pt[1] + i * hCell, k.response))
    return np.array(out, np.float32).reshape(-1, 3)


def extract(P, image):
    """ORBextractor::operator(), ORBextractor.cc:1043-1105, on cv2 primitives."""
    image = np.ascontiguousarray(image, np.uint8)
    if image.size == 0:
        return _empty_result(P.nlevels)
    levels = compute_pyramid(P, image)
    xs, ys, angs, resps, octs, sizes, descs = [], [], [], [], [], [], []
    level_counts = np.zeros(P.nlevels, np.int32)
    minB = EDGE_THRESHOLD - 3
    for l, img in enumerate(levels):
        h, w = img.shape
        maxBX, maxBY = w - EDGE_THRESHOLD + 3, h - EDGE_THRESHOLD + 3
        if maxBX - minB < 30 or maxBY - minB < 30:
            continue
        cands = detect_cells_cv2(img, P.ini_th, P.min_th)
        sel = distribute_octree(cands, minB, maxBX, minB, maxBY, P.features_per_level[l])
        level_counts[l] = len(sel)
        if len(sel) == 0:
            continue
        blur = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
        scale = np.float32(P.scale[l])
        patch = float(int(np.float32(PATCH_SIZE) * scale))
        for s in sel:
            px, py = np.float32(cands[s, 0] + minB), np.float32(cands[s, 1] + minB)
            ix, iy = int(np.rint(px)), int(np.rint(py))
            m01, m10 = ic_moments(P, img, ix, iy)
            ang = np.float32(cv2.fastAtan2(float(m01), float(m10)))
            descs.append(orb_descriptor(blur, ix, iy, ang))
            xs.append(px * scale if l else px)
            ys.append(py * scale if l else py)
            angs.append(ang); resps.append(cands[s, 2]); octs.append(l); sizes.append(patch)
    n = len(xs)
    return dict(x=np.array(xs, np.float32), y=np.array(ys, np.float32), angle=np.array(angs, np.float32),
                response=np.array(resps, np.float32), octave=np.array(octs, np.int32),
                size=np.array(sizes, np.float32),
                desc=np.array(descs, np.uint8).reshape(n, 32), level_counts=level_counts)
