"""TEST INFRASTRUCTURE -- CPU restatement of the reference's training-mask makers.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU legs may import this; the product path never does.

2D: generate_paths.visualize_path (generate_paths.py:118-169), the lines that produce the saved training mask
    (`path_brightness`; the lines tagged VISUAL there only make the preview image):
        path_white = union of cv2.circle(center = path point, radius 10, filled)          :129-131
        path_white = cv2.GaussianBlur(path_white, (33, 33), cv2.BORDER_WRAP)              :134  (3rd positional = sigmaX = 3)
        path_white[point] = 255 for every path point                                      :136-137
        path_brightness = path_white where the map is > 240 (free), else 0               :148,155
    The two OpenCV calls are third-party code that is not in /root/reference (opencv-python; 4.13.0 in this image), so their
    published algorithms are restated here and pinned against cv2 itself (tests/test_mask_oracle.py, run where cv2 exists)
    and against masks produced by the unmodified reference function (tests/golden/masks2d.npz):
      * cv::circle(thickness < 0, LINE_8, shift 0) = Circle() of imgproc/drawing.cpp: the midpoint circle walk, every
        visited (dx, dy) fills rows cy +- dy with half-width dx and rows cy +- dx with half-width dy;
      * cv::GaussianBlur on CV_8U = the fixed-point path of imgproc/smooth.dispatch.cpp: the normalised kernel in 8.8
        fixed point with the rounding error diffused from the ends to the centre (getGaussianKernelFixedPoint_ED), a
        horizontal pass in 16-bit, a vertical pass in 32-bit, result (v + 2^15) >> 16, borders BORDER_REFLECT_101.
3D: ThreeD_enlarge_path.enlarge_the_path (ThreeD_enlarge_path.py:25-65): every path voxel paints the free, in-range
    voxels of the Chebyshev shells i = 0..LAYERS-1 around it with 255 - (i+1) (255/LAYERS - 2) where that is larger than
    what is there (pure NumPy: pinned against the unmodified function, tests/golden/masks3d.npz).
"""
import numpy as np

RADIUS = 10          # generate_paths.py:131
KSIZE = 33           # generate_paths.py:134
SIGMA = 3.0          # cv2.BORDER_WRAP == 3 passed as sigmaX
FREE_ABOVE = 240     # generate_paths.py:148
LAYERS = 5           # ThreeD_enlarge_path.py:4


def disk_half_widths(radius=RADIUS):
    """half_width[|dy|] of the filled circle cv::circle draws (midpoint walk of drawing.cpp Circle())."""
    hw = np.full(radius + 1, -1, np.int64)
    err, dx, dy, plus, minus = 0, radius, 0, 1, (radius << 1) - 1
    while dx >= dy:
        hw[dy] = max(hw[dy], dx)
        hw[dx] = max(hw[dx], dy)
        dy += 1
        err += plus
        plus += 2
        mask = (1 if err <= 0 else 0) - 1     # 0 or -1
        err -= minus & mask
        dx += mask
        minus -= mask & 2
    return hw


def gaussian_kernel_fixed(n=KSIZE, sigma=SIGMA, bits=8):
    """Fixed-point Gaussian taps of cv::GaussianBlur for 8-bit images; they sum to 2^bits exactly."""
    x = np.arange(n) - (n - 1) // 2
    k = np.exp(-0.5 / (sigma * sigma) * x * x)
    k = k / k.sum()
    res = np.zeros(n, np.int64)
    err, s = 0.0, 0
    for i in range(n // 2):
        adj = k[i] * (1 << bits) + err
        v0 = int(np.rint(adj))
        err = adj - v0
        res[i] = res[n - 1 - i] = v0
        s += v0
    res[n // 2] = (1 << bits) - 2 * s
    return res


def draw_disks(shape, path, radius=RADIUS):
    img = np.zeros(shape, np.uint8)
    hw = disk_half_widths(radius)
    H, W = shape
    for (py, px) in np.asarray(path).reshape(-1, 2):
        for dy in range(-radius, radius + 1):
            y = py + dy
            if 0 <= y < H:
                w = hw[abs(dy)]
                img[y, max(px - w, 0):max(min(px + w, W - 1) + 1, 0)] = 255
    return img


def gaussian_blur_u8(src, n=KSIZE, sigma=SIGMA):
    kf = gaussian_kernel_fixed(n, sigma)
    r = n // 2
    H, W = src.shape
    p = np.pad(src.astype(np.int64), r, mode="reflect")          # BORDER_REFLECT_101
    h = np.zeros((H + 2 * r, W), np.int64)
    for i in range(n):
        h += kf[i] * p[:, i:i + W]
    v = np.zeros((H, W), np.int64)
    for j in range(n):
        v += kf[j] * h[j:j + H]
    return np.clip((v + 32768) >> 16, 0, 255).astype(np.uint8)


def mask_2d(occ, path):
    """The image visualize_path() writes to paths/ for one task; path = [(y, x), ...] as astar() returns it."""
    path = np.asarray(path).reshape(-1, 2)
    img = gaussian_blur_u8(draw_disks(occ.shape, path))
    img[path[:, 0], path[:, 1]] = 255
    return np.where(occ > FREE_ABOVE, img, 0).astype(np.uint8)


def layer_values(layers=LAYERS):
    """Value painted on shell i: Python float arithmetic of ThreeD_enlarge_path.py:44-50, stored into a uint8 array."""
    vals, v = [], 255
    for _ in range(layers):
        v -= 255 / layers - 2
        vals.append(np.array([v]).astype(np.uint8)[0])
    return np.array(vals, np.uint8)


def enlarge_3d(occ, path_vol, layers=LAYERS):
    """enlarge_the_path on arrays: occ holds 255 = obstacle; path_vol holds 255 on the path."""
    vals = layer_values(layers)
    on = path_vol == 255
    X, Y, Z = occ.shape
    r = layers - 1
    best = np.zeros(occ.shape, np.uint8)
    pad = np.pad(on, r)
    for dx in range(-r, r + 1):
        for dy in range(-r, r + 1):
            for dz in range(-r, r + 1):
                i = max(abs(dx), abs(dy), abs(dz))
                sh = pad[r + dx:r + dx + X, r + dy:r + dy + Y, r + dz:r + dz + Z]
                best = np.maximum(best, np.where(sh, vals[i], 0).astype(np.uint8))
    out = np.array(path_vol)
    upd = (occ != 255) & (best > out)
    out[upd] = best[upd]
    return out


def path_volume(shape, path):
    """ThreeD_generate_paths.save_and_visualize_path:142-150: zeros with 255 on the path voxels."""
    vol = np.zeros(shape, np.uint8)
    p = np.asarray(path).reshape(-1, 3)
    vol[p[:, 0], p[:, 1], p[:, 2]] = 255
    return vol
