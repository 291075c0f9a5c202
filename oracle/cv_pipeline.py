"""The reference's OpenCV pipeline, called through the cv2 Python wheel.

TEST INFRASTRUCTURE ONLY.  This is the strongest pin available in this image for
`_get_cost` (dpose_core.cpp:55-114): the same OpenCV functions the reference calls, in
the same order, through cv2 (4.13.0 here; ROS noetic ships 4.2).  IPP is switched off
because the distro build the reference links against has no IPP and the IPP
distanceTransform differs by 1 ulp on some pixels.
"""
import numpy as np


def cv_cost_data(footprint, padding):
    """Returns dict(cost, pre_blur, edt, outline, mask, center, box) like cost_data."""
    import cv2
    cv2.ipp.setUseIPP(False)
    cv2.setNumThreads(1)
    fp = np.asarray(footprint, dtype=np.int32).reshape(-1, 2)
    pad = int(padding)
    # dpose_core.cpp:124-130
    center = pad - fp.min(axis=0)
    pts = (fp + center).astype(np.int32)
    # :61-69
    x, y, w, h = cv2.boundingRect(pts.reshape(-1, 1, 2))
    cols, rows = w + 2 * pad, h + 2 * pad
    # :72-73
    image = np.full((rows, cols), 255, dtype=np.uint8)
    cv2.polylines(image, [pts.reshape(-1, 1, 2)], True, 0)
    outline = image.copy()
    # :81-83
    edt = cv2.distanceTransform(image, cv2.DIST_L2, cv2.DIST_MASK_PRECISE, dstType=cv2.CV_32F)
    edt0 = edt.copy()
    # :92-94
    image[:] = 255
    cv2.fillPoly(image, [pts.reshape(-1, 1, 2)], 0)
    mask = image.copy()
    # :100-101 copyTo(edt, outer, image)
    outer = np.zeros((rows, cols), dtype=np.float32)
    outer = cv2.copyTo(edt, image, outer)
    # :104 `outer *= -0.01` is Mat::convertTo(outer, type, -0.01): float alpha, fl32(x*a)
    outer = (outer * np.float32(-0.01)).astype(np.float32)
    # :105-106
    mn, _, _, _ = cv2.minMaxLoc(outer)
    # :109-110
    edt = cv2.copyTo(outer, image, edt)
    edt = cv2.subtract(edt, float(mn))
    pre_blur = edt.copy()
    # :111
    cost = cv2.GaussianBlur(edt, (3, 3), 0)
    # :138
    ring = np.array([[0, 0], [cols, 0], [cols, rows], [0, rows], [0, 0]], dtype=np.int32)
    return dict(cost=cost, pre_blur=pre_blur, edt=edt0, outline=outline, mask=mask,
                center=center.astype(np.int32), box=ring - center, min_outer=float(mn))


def np_get_cost(cost, center, se2, cells, want_jacobian=True):
    """float64 numpy restatement of pose_gradient::get_cost (dpose_core.cpp:200-260),
    independent of the C restatement (vectorised, same FD Jacobian)."""
    cost = np.asarray(cost, dtype=np.float32)
    rows, cols = cost.shape
    cells = np.asarray(cells, dtype=np.float64).reshape(-1, 2)
    cx, cy = float(center[0]), float(center[1])

    def k_of(x, y, th):
        c, s = np.cos(th), np.sin(th)
        # m_to_k = T(x,y) R(th) T(-c); inverse = (R^T, -R^T t)
        t0 = (c * -cx + -s * -cy) + x
        t1 = (s * -cx + c * -cy) + y
        i0 = -(c * t0 + s * t1)
        i1 = -(-s * t0 + c * t1)
        return (c * cells[:, 0] + s * cells[:, 1]) + i0, (-s * cells[:, 0] + c * cells[:, 1]) + i1

    def rnd(v):  # std::round, half away from zero
        return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5))

    kx, ky = k_of(se2[0], se2[1], se2[2])
    ux, uy = rnd(kx).astype(np.int64), rnd(ky).astype(np.int64)
    ok = (ux - 1 >= 1) & (uy - 1 >= 1) & (ux < cols - 1) & (uy < rows - 1)
    ux, uy = ux[ok], uy[ok]
    m00 = cost[uy - 1, ux - 1].astype(np.float64)
    m01 = cost[uy, ux - 1].astype(np.float64)
    m10 = cost[uy - 1, ux].astype(np.float64)
    m11 = cost[uy, ux].astype(np.float64)

    def get(kx, ky):
        rx = kx[ok] - ux + 0.5
        ry = ky[ok] - uy + 0.5
        r0 = (1 - rx) * m00 + rx * m10
        r1 = (1 - rx) * m01 + rx * m11
        return r0 * (1 - ry) + r1 * ry

    total = float(np.sum(get(kx, ky)))
    J = np.zeros(3)
    if want_jacobian:
        h = 1e-6
        for a in range(3):
            d = np.zeros(3)
            d[a] = h
            lo = get(*k_of(se2[0] - d[0], se2[1] - d[1], se2[2] - d[2]))
            hi = get(*k_of(se2[0] + d[0], se2[1] + d[1], se2[2] + d[2]))
            J[a] = float(np.sum((hi - lo) * 5e5))
    return total, J
