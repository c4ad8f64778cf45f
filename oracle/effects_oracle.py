"""CPU oracle for the depth-aware effects (TEST INFRASTRUCTURE ONLY -- imported by tests/, never by the product path).

numpy restatement of the depth-dependent parts of the reference's src/post_processor.py, following its float32
evaluation order.  The two OpenCV calls the reference makes on the depth map are restated from OpenCV's published
algorithms (opencv-python is an un-vendored dependency of the reference, README.md:80; 4.13 in this container) and are
pinned against cv2 itself by tests/test_effects_oracle.py when cv2 is importable:
  * cv2.Sobel(src, CV_32F, dx, dy, ksize=3): separable (-1,0,1) x (1,2,1), BORDER_REFLECT_101;
  * cv2.bilateralFilter(src32f, 9, 75, 75): disc of radius 4, w = exp(-r^2/(2 ss^2)) exp(-dI^2/(2 sc^2)), REFLECT_101
    (cv2 interpolates the colour term from a 4096-bin table: agreement ~1e-6, not bitwise).
Random draws of the reference (hologram noise, interference columns) are arguments here."""
import numpy as np


def normalise_depth(depth):
    """:64-66 / :405-408 / :473-477."""
    d = np.array(depth, dtype=np.float32, copy=True)
    if d.ndim > 2:
        d = d[:, :, 0]
    if d.max() > 1.0:
        d = d / d.max()
    return d


def fog(image, depth, fog_start=0.0, power=3.0, visibility=0.3, fog_color=(255, 255, 255)):
    """_effect_fog, src/post_processor.py:451-493 (power 3.0 and visibility 0.3 are literals there)."""
    color = np.array(fog_color, dtype=np.float32)
    d = normalise_depth(depth)
    a = np.maximum(d - fog_start, 0.0) / (1.0 - fog_start)
    a = np.clip(a, 0.0, 1.0)
    a = a ** power
    a = a * visibility
    f3 = np.stack([a] * 3, axis=2)
    result = image.astype(np.float32) * f3 + color * (1.0 - f3)
    return np.clip(result, 0, 255).astype(np.uint8)


def _pad101(a, r):
    return np.pad(a, r, mode="reflect")   # numpy 'reflect' == BORDER_REFLECT_101


def sobel_magnitude(src):
    """sqrt(Sobel_x^2 + Sobel_y^2), :72-74 / :414-416."""
    p = _pad101(src.astype(np.float32), 1)
    l, c, r = p[:, :-2], p[:, 1:-1], p[:, 2:]
    dx_rows = r - l                               # derivative along x, per row
    sx_rows = (l + r) + c * np.float32(2.0)       # smoothing along x, per row
    gx = (dx_rows[:-2] + dx_rows[2:]) + dx_rows[1:-1] * np.float32(2.0)
    gy = sx_rows[2:] - sx_rows[:-2]
    return np.sqrt(gx * gx + gy * gy)


def bilateral(src, d=9, sigma_color=75.0, sigma_space=75.0):
    """cv2.bilateralFilter on float32 (:69)."""
    src = src.astype(np.float32)
    radius = d // 2
    p = _pad101(src, radius)
    h, w = src.shape
    num = np.zeros((h, w), np.float32)
    den = np.zeros((h, w), np.float32)
    sc = np.float32(-0.5 / (sigma_space * sigma_space))
    cc = np.float32(-0.5 / (sigma_color * sigma_color))
    for dy in range(-radius, radius + 1):
        for dx in range(-radius, radius + 1):
            r2 = dx * dx + dy * dy
            if r2 > radius * radius:
                continue
            v = p[radius + dy: radius + dy + h, radius + dx: radius + dx + w]
            dv = v - src
            wgt = np.exp(np.float32(r2) * sc) * np.exp(dv * dv * cc)
            num += v * wgt
            den += wgt
    return num / den


def edge_mask(mag):
    """:77-85: normalise by the maximum, threshold 0.05, 3x3 dilation."""
    g = mag
    if g.max() > 0:
        g = g / g.max()
    e = np.where(g > 0.05, 1.0, 0.0).astype(np.float32)
    p = np.pad(e, 1, mode="constant")
    out = np.zeros_like(e)
    for dy in range(3):
        for dx in range(3):
            out = np.maximum(out, p[dy: dy + e.shape[0], dx: dx + e.shape[1]])
    return out


def toon(image, depth, levels=5, edge_strength=1.0):
    """_effect_toon with depth, src/post_processor.py:64-102."""
    img = image.astype(np.float32)
    q = np.floor(img / 255.0 * levels) / levels * 255.0
    mag = sobel_magnitude(bilateral(normalise_depth(depth)))
    e3 = np.stack([edge_mask(mag)] * 3, axis=2)
    return np.clip(q * (1 - edge_strength * e3), 0, 255).astype(np.uint8), mag


def hologram(image, depth, num_lines=50, noise=None, lines=()):
    """_effect_hologram, src/post_processor.py:373-449; noise (:399) and lines [(x_pos, x_width)] (:443-446) are inputs."""
    img = image.astype(np.float32) / 255.0
    cyan = np.zeros_like(img)
    cyan[:, :, 0] = img[:, :, 0] * 0.8
    cyan[:, :, 1] = img[:, :, 1] * 1.0
    cyan[:, :, 2] = img[:, :, 2] * 0.2
    height, width = image.shape[:2]
    line_height = height / num_lines
    scan = np.ones_like(img)
    for i in range(num_lines):
        y_start = int(i * line_height)
        y_end = int(min((i + 0.7) * line_height, height))
        scan[y_start:y_end, :, :] *= 0.85
    base = cyan * scan
    glow = np.zeros_like(img)
    if depth is not None:
        e = sobel_magnitude(normalise_depth(depth))
        if e.max() > 0:
            e = e / e.max()
        glow = np.stack([e * 0.1, e * 0.6, e * 0.3], axis=2)
    nz = np.zeros_like(img) if noise is None else noise.astype(np.float32)
    holo = base + glow + nz
    for x_pos, x_width in lines:
        holo[:, x_pos:min(x_pos + x_width, width), :] *= 1.5
    return np.clip(holo * 255, 0, 255).astype(np.uint8)
