"""Synthetic inputs of SURVEY.md section 8(d) M2, shared by tests/ and bench.py (test infrastructure).

Hashes: counter-based splitmix64 (index addressable, numpy), seed 0xB200; word 15 keeps 40 random bits
(mirrors VideoHash::random_hash, video_hash.rs:293-306); 10 % of entries are perturbed copies of an earlier
entry.  Frame stacks: sums of 4 spatio-temporal cosines + U[-8,8] noise, with letterbox / pillarbox bars and
static stacks mixed in; built with torch so the same code fills HBM directly on the GPU box.
"""
from __future__ import annotations

import numpy as np

SEED = 0xB200
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """one splitmix64 output per counter value (vectorised, wraps mod 2^64)"""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _stream(seed: int, idx: np.ndarray, lane: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        return splitmix64(splitmix64(np.uint64(seed) ^ (np.uint64(lane) * np.uint64(0xD1342543DE82EF95))) + idx.astype(np.uint64))


def random_hashes(n: int, seed: int = SEED, start: int = 0) -> np.ndarray:
    """[n,16] u64; entry i depends only on (seed, start+i)"""
    idx = np.arange(start, start + n, dtype=np.uint64)
    h = np.stack([_stream(seed, idx, w) for w in range(16)], axis=1)
    h[:, 15] &= np.uint64((1 << 40) - 1)
    return h


def planted_hashes(n: int, seed: int = SEED, dup_frac_den: int = 10, max_flip: int = 300, chunk: int = 1 << 15, pad_flips: bool = False):
    """[n,16] u64 where ~1/dup_frac_den of the entries are copies of the BASE hash of a random earlier entry
    with each of the 1000 hash bits flipped with probability k/1024, k ~ U[0, max_flip].  -> (hashes, src) with
    src[i] = i for base entries.  Bits 1000..1023 stay zero, as in every real VideoHash (dct_3d.rs:55-66 writes 1000
    bits); pad_flips=True flips them too (round 1's generator; the test helpers of video_hash.rs:265-280 make such hashes)."""
    h = random_hashes(n, seed)
    idx = np.arange(n, dtype=np.uint64)
    is_dup = (_stream(seed, idx, 100) % np.uint64(dup_frac_den) == 0) & (idx > 0)
    src = idx.copy()
    d = np.nonzero(is_dup)[0]
    src[d] = _stream(seed, d.astype(np.uint64), 101) % d.astype(np.uint64)
    base = h.copy()
    for a in range(0, len(d), chunk):
        dd = d[a:a + chunk]
        k = (_stream(seed, dd.astype(np.uint64), 102) % np.uint64(max_flip + 1)).astype(np.int64)
        rng = np.random.Generator(np.random.Philox(key=seed + 7, counter=[0, 0, 0, a]))
        u = rng.integers(0, 1024, (len(dd), 1024), dtype=np.int16)
        flip = np.packbits(u < k[:, None], axis=1, bitorder="little").view(np.uint64)
        if not pad_flips:
            flip[:, 15] &= np.uint64((1 << 40) - 1)
        h[dd] = base[src[dd].astype(np.int64)] ^ flip
    return h, src


def lognormal_durations(n: int, seed: int = SEED) -> np.ndarray:
    rng = np.random.Generator(np.random.Philox(key=seed + 11))
    return np.clip(np.exp(rng.normal(np.log(600.0), 1.0, n)), 5, 14400).astype(np.uint32)


def paths(n: int, start: int = 0):
    """fixed-width names under one directory: byte order == Path order"""
    return ["v/%08d" % (start + i) for i in range(n)]


def smooth_frame(w: int, h: int, seed: int) -> np.ndarray:
    """one low-frequency u8 frame (numpy) for resize tests"""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.full((h, w), 128.0)
    for _ in range(4):
        a = rng.uniform(20, 50)
        fx, fy = rng.uniform(-3, 3, 2)
        img += a * np.cos(2 * np.pi * (fx * xx / w + fy * yy / h) + rng.uniform(0, 2 * np.pi))
    img += rng.integers(-8, 9, (h, w))
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def frame_stacks(n: int, w: int, h: int, seed: int = SEED, device="cpu", first_id: int = 0, bars: bool = True):
    """[n,16,h,w] u8 torch tensor on `device`; stack s depends only on (seed, first_id+s) for its parameters
    (the additive noise comes from torch's generator seeded per call).  25 % letterboxed (bars 16+-4, 5-12 %
    of the height), 10 % pillarboxed, 5 % static."""
    import torch

    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + first_id)
    out = torch.empty((n, 16, h, w), dtype=torch.uint8, device=device)
    xs = torch.arange(w, device=device, dtype=torch.float32)[None, None, :] / w
    ys = torch.arange(h, device=device, dtype=torch.float32)[None, :, None] / h
    ts = torch.arange(16, device=device, dtype=torch.float32)[:, None, None] / 16
    for s in range(n):
        prm = np.random.default_rng([seed, first_id + s])
        img = torch.full((16, h, w), 128.0, device=device)
        static = bars and prm.random() < 0.05
        for _ in range(4):
            a = float(prm.uniform(20, 50))
            fx, fy, ft = (float(v) for v in prm.uniform(-3, 3, 3))
            if static:
                ft = 0.0
            ph = float(prm.uniform(0, 2 * np.pi))
            img += a * torch.cos(2 * np.pi * (fx * xs + fy * ys + ft * ts) + ph)
        noise = torch.randint(-8, 9, (1 if static else 16, h, w), device=device, generator=g)
        img = (img + noise).round_().clamp_(0, 255).to(torch.uint8)
        u = prm.random()
        if bars and u < 0.25:
            t, b = (int(prm.uniform(0.05, 0.12) * h) for _ in range(2))
            img[:, :t, :] = (16 + torch.randint(-4, 5, (16, t, w), device=device, generator=g)).to(torch.uint8)
            img[:, h - b:, :] = (16 + torch.randint(-4, 5, (16, b, w), device=device, generator=g)).to(torch.uint8)
        elif bars and u < 0.35:
            l, r = (int(prm.uniform(0.05, 0.12) * w) for _ in range(2))
            img[:, :, :l] = (16 + torch.randint(-4, 5, (16, h, l), device=device, generator=g)).to(torch.uint8)
            img[:, :, w - r:] = (16 + torch.randint(-4, 5, (16, h, r), device=device, generator=g)).to(torch.uint8)
        out[s] = img
    return out


# ---- large tables (10 M hashes): the same kind of data, generated with torch integer arithmetic so that the GPU box fills
# HBM directly; bit-identical on CPU and GPU (wrapping int64 ops only), entry i depends only on (seed, i) ---------------
def _t_splitmix64(x):
    import torch

    def lsr(v, s):  # logical shift right on int64
        return (v >> s) & ((1 << (64 - s)) - 1)

    z = x + (-7046029254386353131)  # 0x9E3779B97F4A7C15 as int64
    z = (z ^ lsr(z, 30)) * (-4658895280553007687)  # 0xBF58476D1CE4E5B9
    z = (z ^ lsr(z, 27)) * (-7723592293110705685)  # 0x94D049BB133111EB
    return z ^ lsr(z, 31)


def _t_stream(seed: int, idx, lane: int):
    import torch

    k = torch.tensor([(seed ^ (lane * 0xD1342543DE82EF95)) & 0xFFFFFFFFFFFFFFFF], dtype=torch.uint64).view(torch.int64).to(idx.device)
    return _t_splitmix64(_t_splitmix64(k) + idx)


def planted_hashes_torch(n: int, seed: int = SEED, device="cpu", dup_frac_den: int = 10, chunk: int = 1 << 21):
    """[n,16] int64 tensor (the u64 words) on `device`: random 1000-bit hashes (bits 1000..1023 zero) where ~1/dup_frac_den
    of the entries are copies of a random EARLIER base entry with a random subset of bits flipped: the AND of m in 2..6
    random words per word, i.e. ~250, 125, 62, 31 or 16 flipped bits.  -> (hashes, src), src[i] = i for base entries."""
    import torch

    out = torch.empty((n, 16), dtype=torch.int64, device=device)
    src = torch.empty(n, dtype=torch.int64, device=device)
    low40 = (1 << 40) - 1

    def base(idx):
        h = torch.stack([_t_stream(seed, idx, w) for w in range(16)], dim=1)
        h[:, 15] &= low40
        return h

    for a in range(0, n, chunk):
        idx = torch.arange(a, min(n, a + chunk), dtype=torch.int64, device=device)
        h = base(idx)
        r = _t_stream(seed, idx, 100) & 0x7FFFFFFFFFFFFFFF
        is_dup = (r % dup_frac_den == 0) & (idx > 0)
        s = torch.where(is_dup, (_t_stream(seed, idx, 101) & 0x7FFFFFFFFFFFFFFF) % torch.clamp(idx, min=1), idx)
        d = torch.nonzero(is_dup).flatten()
        if d.numel():
            di = idx[d]
            hb = base(s[d])
            m = 2 + (_t_stream(seed, di, 102) & 0x7FFFFFFFFFFFFFFF) % 5  # 2..6 words ANDed
            flip = torch.full((d.numel(), 16), -1, dtype=torch.int64, device=device)
            for k in range(6):
                rk = torch.stack([_t_stream(seed, di, 200 + 16 * k + w) for w in range(16)], dim=1)
                flip = torch.where((m > k)[:, None], flip & rk, flip)
            flip[:, 15] &= low40
            h[d] = hb ^ flip
        out[a:a + chunk] = h
        src[a:a + chunk] = s
    return out, src
