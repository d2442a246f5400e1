"""Deterministic (numpy-seeded) inputs shared by the golden generator (oracle/gen_golden_gpu.py) and the tests,
so that CPU tests can regenerate bit-identical inputs for fixtures that store only the reference's outputs."""
import math

import numpy as np


def camera_rays(N, seed=0, radius=3.35, spread=0.16, jitter=0.02):
    """N rays from around (0, radius, 0) looking roughly along -y with some spread (covers hits and misses)."""
    rs = np.random.RandomState(seed)
    o = np.array([0.0, radius, 0.0], np.float32) + (rs.rand(N, 3).astype(np.float32) - 0.5) * jitter
    d = np.stack([(rs.rand(N) - 0.5) * 2 * spread, -np.ones(N), (rs.rand(N) - 0.5) * 2 * spread], -1)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)


def inside_rays(N, seed=0, bound=1.0):
    """rays starting inside the volume with arbitrary directions (all sign combinations, axis-aligned cases)."""
    rs = np.random.RandomState(seed)
    o = ((rs.rand(N, 3) - 0.5) * 1.6 * bound * np.array([1, 0.5, 1])).astype(np.float32)
    d = rs.randn(N, 3)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    return np.ascontiguousarray(o, np.float32), np.ascontiguousarray(d, np.float32)


def random_bitfield(C, H, p=0.3, seed=1):
    rs = np.random.RandomState(seed)
    occ = rs.rand(C * H ** 3) < p
    return np.packbits(occ, bitorder='little')


def full_bitfield(C, H):
    return np.full(C * H ** 3 // 8, 255, np.uint8)


def aabb_of(bound):
    return np.array([-bound, -bound / 2, -bound, bound, bound / 2, bound], np.float32)


def grid_setup(D, L=16, C=2, base=16, log2_hash=16, desired=2048, seed=0):
    from oracle.field import grid_offsets
    offsets, pls = grid_offsets(D, num_levels=L, level_dim=C, base_resolution=base, log2_hashmap_size=log2_hash, desired_resolution=desired)
    rs = np.random.RandomState(seed)
    emb = ((rs.rand(int(offsets[-1]), C) - 0.5)).astype(np.float32)
    return offsets, float(np.float32(np.log2(pls))), emb


def unit_points(B, D, seed=0, with_edges=True):
    rs = np.random.RandomState(seed)
    x = rs.rand(B, D).astype(np.float32)
    if with_edges and B >= 8:
        x[0] = 0.0
        x[1] = 1.0
        x[2] = 0.5
        x[3, 0] = 1.0000001   # out of range -> zeros
        x[4, 0] = -1e-7
        x[5] = np.float32(1.0) - np.float32(2 ** -24)
    return x


def field_samples(M, seed=0, bound=1.0):
    rs = np.random.RandomState(seed)
    xyz = ((rs.rand(M, 3) * 2 - 1) * bound * np.array([1, 0.5, 1])).astype(np.float32)
    d = rs.randn(M, 3)
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    return xyz, d
