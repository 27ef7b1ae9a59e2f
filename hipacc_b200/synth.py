"""Counter-based synthetic images.

    v(x, y, seed) = splitmix64(seed XOR (y * 2^32 + x))
    uchar pixel   = v & 0xFF                (uniform 0..255, like rand() % 256,
                                             samples-public/common/hipacc_helper.hpp:87)
    float pixel   = (v >> 40) * 2^-24       (uniform [0,1), like rand()/RAND_MAX, :85)

Being a pure function of (x, y, seed), any row strip can be produced on any rank / device and
equals the corresponding slice of the full image.  Two implementations with identical results:
numpy (uint64) for hosts and torch int64 (wrapping multiply, masked logical shifts) for
generating directly in HBM.
"""
import numpy as np

_M64 = (1 << 64) - 1
_C0 = 0x9E3779B97F4A7C15
_C1 = 0xBF58476D1CE4E5B9
_C2 = 0x94D049BB133111EB


def _splitmix64_np(z):
    with np.errstate(over="ignore"):
        z = z + np.uint64(_C0)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_C1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_C2)
        return z ^ (z >> np.uint64(31))


def _bits_np(width, height, seed, x0=0, y0=0):
    x = np.arange(x0, x0 + width, dtype=np.uint64)[None, :]
    y = np.arange(y0, y0 + height, dtype=np.uint64)[:, None]
    return _splitmix64_np(np.uint64(seed) ^ ((y << np.uint64(32)) + x))


def image_np(dtype, width, height, seed=1, x0=0, y0=0, scale=1.0):
    """dtype 'uint8' | 'float32' | 'int8'.  `scale` multiplies float pixels (C3 uses 255)."""
    v = _bits_np(width, height, seed, x0, y0)
    if dtype == "uint8":
        return (v & np.uint64(0xFF)).astype(np.uint8)
    if dtype == "int8":
        return (v & np.uint64(0xFF)).astype(np.uint8).view(np.int8)
    if dtype == "float32":
        f = ((v >> np.uint64(40)).astype(np.float32)) * np.float32(2.0 ** -24)
        return f if scale == 1.0 else (f * np.float32(scale)).astype(np.float32)
    raise ValueError(dtype)


def _i64(c):
    return c - (1 << 64) if c >= (1 << 63) else c


def image_torch(dtype, width, height, seed=1, x0=0, y0=0, scale=1.0, device="cpu", rows_per_chunk=2048):
    """Same pixels as image_np, generated with torch int64 ops on `device` (e.g. 'cuda')."""
    import torch
    tdt = {"uint8": torch.uint8, "float32": torch.float32}[dtype]
    out = torch.empty((height, width), dtype=tdt, device=device)
    x = torch.arange(x0, x0 + width, dtype=torch.int64, device=device)[None, :]

    def lsr(z, s):  # logical shift right on two's-complement int64
        return (z >> s) & ((1 << (64 - s)) - 1)

    for r0 in range(0, height, rows_per_chunk):
        r1 = min(height, r0 + rows_per_chunk)
        y = torch.arange(y0 + r0, y0 + r1, dtype=torch.int64, device=device)[:, None]
        z = (y << 32) + x
        z = z ^ _i64(seed & _M64)
        z = z + _i64(_C0)
        z = (z ^ lsr(z, 30)) * _i64(_C1)
        z = (z ^ lsr(z, 27)) * _i64(_C2)
        z = z ^ lsr(z, 31)
        if dtype == "uint8":
            out[r0:r1] = (z & 0xFF).to(torch.uint8)
        else:
            f = lsr(z, 40).to(torch.float32) * (2.0 ** -24)
            out[r0:r1] = f if scale == 1.0 else f * scale
    return out


def blocks_np(width, height, seed=1, n_rect=40):
    """Piecewise-constant uchar test image (random rectangles + low noise): gives sparse, well
    defined Harris corners, unlike white noise."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 40, dtype=np.int32)
    for _ in range(n_rect):
        w = int(rng.integers(4, max(5, width // 3)))
        h = int(rng.integers(4, max(5, height // 3)))
        x = int(rng.integers(0, max(1, width - w)))
        y = int(rng.integers(0, max(1, height - h)))
        img[y:y + h, x:x + w] = int(rng.integers(0, 256))
    img += rng.integers(-2, 3, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)
