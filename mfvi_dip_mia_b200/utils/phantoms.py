"""Synthetic inputs for the BASELINE configs (the reference's real images are not shipped;
SURVEY.md §8d).  Pure numpy, analytic, no RNG except where a seed is passed explicitly."""
import numpy as np

# Shepp-Logan ellipses: (intensity, a, b, x0, y0, phi_deg) — the standard modified phantom.
_SHEPP_LOGAN = [
    (1.0, .69, .92, 0, 0, 0), (-.8, .6624, .8740, 0, -.0184, 0), (-.2, .1100, .3100, .22, 0, -18),
    (-.2, .1600, .4100, -.22, 0, 18), (.1, .2100, .2500, 0, .35, 0), (.1, .0460, .0460, 0, .1, 0),
    (.1, .0460, .0460, 0, -.1, 0), (.1, .0460, .0230, -.08, -.605, 0), (.1, .0230, .0230, 0, -.606, 0),
    (.1, .0230, .0460, .06, -.605, 0),
]
_ELLIPSES = [
    (0.85, .80, .60, 0.0, 0.0, 20), (-0.35, .35, .22, -.25, .12, -30), (0.30, .18, .30, .30, -.15, 45),
    (-0.25, .10, .10, .05, .35, 0), (0.40, .07, .16, -.40, -.30, 70), (0.20, .25, .08, .10, -.45, 10),
]


def _ellipses(n, table):
    y, x = np.mgrid[-1:1:n * 1j, -1:1:n * 1j]
    img = np.zeros((n, n), np.float64)
    for val, a, b, x0, y0, phi in table:
        p = np.deg2rad(phi)
        xr = (x - x0) * np.cos(p) + (y - y0) * np.sin(p)
        yr = -(x - x0) * np.sin(p) + (y - y0) * np.cos(p)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1.0] += val
    return np.clip(img, 0.0, 1.0).astype(np.float32)


def ellipse_phantom(n=256):
    """(1,n,n) float32 in [0,1] — denoising / SR target."""
    return _ellipses(n, _ELLIPSES)[None]


def shepp_logan(n=512):
    """(1,n,n) float32 in [0,1] — CT target."""
    return _ellipses(n, _SHEPP_LOGAN)[None, ::-1].copy()


def rgb_phantom(n=512):
    """(3,n,n) float32 — inpainting target (three shifted ellipse phantoms)."""
    base = _ellipses(n, _ELLIPSES)
    return np.stack([base, np.roll(base, n // 16, 0) * 0.8, np.roll(base, -n // 16, 1) * 0.6 + 0.2 * base]).astype(np.float32)


def random_mask(n=512, seed=2, keep=0.5):
    """(1,n,n) {0,1} float32 Bernoulli(keep) mask — inpainting."""
    rng = np.random.RandomState(seed)
    return (rng.rand(1, n, n) < keep).astype(np.float32)


def noisy(img, sigma=0.1, seed=1):
    """clip(img + N(0,sigma^2)) as utils/denoising_utils.py:11 of the reference does (np.random.seed(seed))."""
    rng = np.random.RandomState(seed)
    return np.clip(img + rng.normal(scale=sigma, size=img.shape), 0, 1).astype(np.float32)
