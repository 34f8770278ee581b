"""Seeded synthetic descriptor sets of the shapes BASELINE.json names.

The reference produces its descriptors with OpenCV detectors (getFeature,
/root/reference/src/Sfm.cpp:303-392): SIFT -> CV_32F x 128, AKAZE-MLDB -> CV_8U x 61
(486 bit), ORB -> CV_8U x 32.  These generators only reproduce the *shapes* and a
match structure (half of every image's rows are noisy views of a shared "world" pool,
so a fraction of queries passes the ratio test and exact distance ties are frequent).
numpy only; no GPU, no oracle.
"""
from __future__ import annotations

import numpy as np

AKAZE_BITS = 486  # 61-byte MLDB descriptor, 2 pad bits in the last byte
ORB_BITS = 256


def binary_world(n_desc: int, bits: int = AKAZE_BITS, seed: int = 0) -> np.ndarray:
    """Shared pool of 2*n_desc bit rows (unpacked, one byte per bit)."""
    return np.random.default_rng(seed).integers(0, 2, (2 * n_desc, bits), dtype=np.uint8)


def binary_image(world: np.ndarray, i: int, n_desc: int, seed: int = 0, flip_p: float = 0.05) -> np.ndarray:
    """Descriptor set of image `i`: (n_desc, ceil(bits/8)) uint8, pad bits zero."""
    bits = world.shape[1]
    rng = np.random.default_rng([seed, i])
    n_shared = n_desc // 2
    pick = rng.choice(world.shape[0], n_shared, replace=False)
    shared = world[pick] ^ (rng.random((n_shared, bits), dtype=np.float32) < flip_p).astype(np.uint8)
    rest = rng.integers(0, 2, (n_desc - n_shared, bits), dtype=np.uint8)
    rows = np.concatenate([shared, rest], 0)
    rng.shuffle(rows, axis=0)
    return np.packbits(rows, axis=1)


_WORLD_CACHE: dict = {}


def binary_image_task(args) -> np.ndarray:
    """binary_image for process pools: args = (i, n_desc, bits, seed); the world is built once per worker."""
    i, n_desc, bits, seed = args
    key = (n_desc, bits, seed)
    if key not in _WORLD_CACHE:
        _WORLD_CACHE.clear()
        _WORLD_CACHE[key] = binary_world(n_desc, bits, seed)
    return binary_image(_WORLD_CACHE[key], i, n_desc, seed)


def binary_images(n_images: int, n_desc, bits: int = AKAZE_BITS, seed: int = 0, flip_p: float = 0.05):
    """List of `n_images` descriptor sets; `n_desc` is an int or a per-image sequence (ragged)."""
    counts = [int(n_desc)] * n_images if np.isscalar(n_desc) else [int(c) for c in n_desc]
    world = binary_world(max(max(counts), 1), bits, seed)
    return [binary_image(world, i, c, seed, flip_p) for i, c in enumerate(counts)]


def _siftify(x: np.ndarray, integer: bool) -> np.ndarray:
    """|x| -> L2 normalise -> clip 0.2 -> renormalise -> x512 -> (floor) -> clip 255: SIFT's own tail."""
    x = np.abs(x).astype(np.float32)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    x = np.minimum(x, 0.2)
    x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
    x = x * 512.0
    if integer:
        x = np.floor(x)
    return np.minimum(x, 255.0).astype(np.float32)


def float_world(n_desc: int, dim: int = 128, seed: int = 0, integer: bool = True) -> np.ndarray:
    return _siftify(np.random.default_rng(seed).standard_normal((2 * n_desc, dim)), integer)


def float_image(world: np.ndarray, i: int, n_desc: int, seed: int = 0, integer: bool = True) -> np.ndarray:
    """SIFT-shape set: (n_desc, dim) float32, non-negative; integer-valued 0..255 when `integer`."""
    dim = world.shape[1]
    rng = np.random.default_rng([seed, i])
    n_shared = n_desc // 2
    pick = rng.choice(world.shape[0], n_shared, replace=False)
    if integer:
        noise = rng.integers(-3, 4, (n_shared, dim)).astype(np.float32)
    else:
        noise = (rng.standard_normal((n_shared, dim)) * 1.5).astype(np.float32)
    shared = np.clip(world[pick] + noise, 0.0, 255.0).astype(np.float32)
    rest = _siftify(rng.standard_normal((n_desc - n_shared, dim)), integer)
    rows = np.concatenate([shared, rest], 0)
    rng.shuffle(rows, axis=0)
    return np.ascontiguousarray(rows)


def float_images(n_images: int, n_desc, dim: int = 128, seed: int = 0, integer: bool = True):
    counts = [int(n_desc)] * n_images if np.isscalar(n_desc) else [int(c) for c in n_desc]
    world = float_world(max(max(counts), 1), dim, seed, integer)
    return [float_image(world, i, c, seed, integer) for i, c in enumerate(counts)]


def all_pairs(n_images: int):
    """findBestPair's enumeration (/root/reference/src/Sfm.cpp:511-512): q<t, row-major."""
    return [(q, t) for q in range(n_images - 1) for t in range(q + 1, n_images)]
