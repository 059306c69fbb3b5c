"""Deterministic synthetic video frames (BGR uint8, HWC) for parity tests and benchmarks.

A smooth court-like background, four moving bright rectangles ("players") and a small moving disc ("ball"), plus
low-amplitude per-pixel noise so resamplers see realistic content.  Pure torch ops: the same code runs
on CPU (tests) and on the GPU (bench); integer arithmetic only, so both produce identical bytes.
"""
from __future__ import annotations

import torch


def _hash_u8(idx: torch.Tensor, seed: int) -> torch.Tensor:
    """Cheap integer hash -> [0, 255] (int64 math, identical on CPU and CUDA)."""
    x = (idx + seed * 7919) & 0xFFFFFFFF
    x = (x ^ (x >> 15)) * 0x2C1B3C6D & 0xFFFFFFFF
    x = (x ^ (x >> 12)) * 0x297A2D39 & 0xFFFFFFFF
    x = x ^ (x >> 15)
    return x & 0xFF


def make_frames(n: int, height: int = 1080, width: int = 1920, seed: int = 1234, start: int = 0,
                device: str | torch.device = "cpu") -> torch.Tensor:
    """(n, height, width, 3) uint8 BGR frames start..start+n-1."""
    dev = torch.device(device)
    ys = torch.arange(height, device=dev, dtype=torch.int64).view(1, height, 1)
    xs = torch.arange(width, device=dev, dtype=torch.int64).view(1, 1, width)
    t = (torch.arange(n, device=dev, dtype=torch.int64) + start).view(n, 1, 1)
    # background: vertical gradient + court lines (integer math)
    base_g = 60 + (ys * 90) // height
    base_b = 90 + (xs * 40) // width
    base_r = 50 + ((xs + ys) * 30) // (width + height)
    line = ((xs * 12 // width) * width // 12 - xs).abs() < max(2, width // 640)
    line = line | (((ys * 6 // height) * height // 6 - ys).abs() < max(2, height // 360))
    noise = _hash_u8(ys * width + xs + t * 0, seed) % 7  # static texture
    frames = torch.empty((n, height, width, 3), dtype=torch.uint8, device=dev)
    chans = []
    for base in (base_b, base_g, base_r):
        c = (base + noise).expand(n, height, width).clone()
        c = torch.where(line.expand(n, height, width), c + 110, c)
        chans.append(c)
    # players: 4 rectangles following slow integer trajectories
    for p in range(4):
        pw, ph = width // 24 + p * (width // 200), height // 6 + p * (height // 100)
        cx = (width * (2 + 2 * p) // 10 + ((t * (3 + p)) % (width // 8)) - width // 16)
        cy = (height * (3 + (p % 2) * 4) // 10 + ((t * (2 + p)) % (height // 10)))
        inside = ((xs - cx).abs() < pw // 2) & ((ys - cy).abs() < ph // 2)
        col = (200 - 30 * p, 60 + 40 * p, 220 - 20 * p)
        for k in range(3):
            chans[k] = torch.where(inside, torch.full_like(chans[k], col[k]) + noise, chans[k])
    # ball: disc of radius ~ width/320
    r = max(3, width // 320)
    bx = (width // 5 + (t * (width // 97)) % (3 * width // 5))
    by = (height // 4 + ((t * (height // 61)) % (height // 2)))
    ball = ((xs - bx) ** 2 + (ys - by) ** 2) <= r * r
    for k, v in enumerate((90, 250, 240)):
        chans[k] = torch.where(ball, torch.full_like(chans[k], v), chans[k])
    for k in range(3):
        frames[..., k] = chans[k].clamp(0, 255).to(torch.uint8)
    return frames


def make_median(height: int = 1080, width: int = 1920, seed: int = 1234, device="cpu") -> torch.Tensor:
    """Background-only RGB uint8 (H,W,3): what a per-pixel median over many frames converges to."""
    dev = torch.device(device)
    ys = torch.arange(height, device=dev, dtype=torch.int64).view(height, 1)
    xs = torch.arange(width, device=dev, dtype=torch.int64).view(1, width)
    base_g = 60 + (ys * 90) // height
    base_b = 90 + (xs * 40) // width
    base_r = 50 + ((xs + ys) * 30) // (width + height)
    line = ((xs * 12 // width) * width // 12 - xs).abs() < max(2, width // 640)
    line = line | (((ys * 6 // height) * height // 6 - ys).abs() < max(2, height // 360))
    noise = _hash_u8(ys * width + xs, seed) % 7
    out = torch.empty((height, width, 3), dtype=torch.uint8, device=dev)
    for k, base in enumerate((base_r, base_g, base_b)):  # RGB order
        c = (base + noise).expand(height, width).clone()
        c = torch.where(line, c + 110, c)
        out[..., k] = c.clamp(0, 255).to(torch.uint8)
    return out
