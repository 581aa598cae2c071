"""Synthetic inputs for parity tests and benchmarks.

All generators are pure integer arithmetic on the byte index, so any slice of
any stream can be produced independently (on the host with numpy here, or on a
GPU with the same formulas in ``torch``): a rank that owns packets [p0, p1) of
a 16 GiB input generates exactly its own bytes.

    word_i     = splitmix64(seed + (i+1) * GOLDEN)            i >= 0
    uniform    = little-endian bytes of word_0, word_1, ...
    and3       = uniform(seed) & uniform(seed+1) & uniform(seed+2)      (~4.35 bits/byte)
    mixed      = 64 KiB regions cycling uniform / and3 / zeros / and2
"""
from __future__ import annotations

import numpy as np

_GOLDEN = 0x9E3779B97F4A7C15
_M1 = 0xBF58476D1CE4E5B9
_M2 = 0x94D049BB133111EB
_MASK = (1 << 64) - 1


def _mix_np(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * np.uint64(_M1)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(_M2)
    return z ^ (z >> np.uint64(31))


def uniform(seed: int, n: int, start: int = 0) -> np.ndarray:
    """Bytes [start, start+n) of the splitmix64 stream ``seed``."""
    if n <= 0:
        return np.zeros(0, dtype=np.uint8)
    w0 = start // 8
    w1 = (start + n + 7) // 8
    with np.errstate(over="ignore"):
        idx = np.arange(w0 + 1, w1 + 1, dtype=np.uint64)
        z = np.uint64(seed & _MASK) + idx * np.uint64(_GOLDEN)
        words = _mix_np(z)
    raw = words.astype("<u8").view(np.uint8)
    return raw[start - 8 * w0: start - 8 * w0 + n].copy()


def and3(seed: int, n: int, start: int = 0) -> np.ndarray:
    return uniform(seed, n, start) & uniform(seed + 1, n, start) & uniform(seed + 2, n, start)


def and2(seed: int, n: int, start: int = 0) -> np.ndarray:
    return uniform(seed, n, start) & uniform(seed + 1, n, start)


def mixed(seed: int, n: int, start: int = 0) -> np.ndarray:
    """64 KiB regions cycling uniform(seed) / and3(seed+10) / zeros / and2(seed+20)."""
    out = np.empty(n, dtype=np.uint8)
    pos = start
    end = start + n
    while pos < end:
        region = pos // 65536
        stop = min(end, (region + 1) * 65536)
        m = stop - pos
        r = region % 4
        if r == 0:
            chunk = uniform(seed, m, pos)
        elif r == 1:
            chunk = and3(seed + 10, m, pos)
        elif r == 2:
            chunk = np.zeros(m, dtype=np.uint8)
        else:
            chunk = and2(seed + 20, m, pos)
        out[pos - start: stop - start] = chunk
        pos = stop
    return out


def zeros(n: int) -> np.ndarray:
    return np.zeros(n, dtype=np.uint8)


def round_robin(n: int) -> np.ndarray:
    """byte j = j & 255 (the reference's worst-ratio regular input, 8281 B/packet)."""
    return (np.arange(n, dtype=np.uint64) & np.uint64(255)).astype(np.uint8)


def adversarial_packet() -> np.ndarray:
    """8192 bytes built greedily: next symbol = highest-index symbol with minimal count."""
    cnt = np.ones(256, dtype=np.int64)
    out = np.empty(8192, dtype=np.uint8)
    for i in range(8192):
        m = cnt.min()
        s = 255 - int(np.argmax(cnt[::-1] == m))
        out[i] = s
        cnt[s] += 1
    return out


def straddle_packet(n: int = 8192, report: bool = False):
    """n bytes built greedily so that the coder's interval keeps containing the midpoint of its window: every
    step takes the symbol whose sub-interval straddles 0x8000 (when there is one), so the reference's
    pending-underflow counter grows into the hundreds -- far beyond the 16 bits the encoders keep pending, which
    is what their rare paths (long underflow runs / a carry into words already stored) need to be exercised.
    Follows the reference coder step (src/gpuar_kernel.cu:256-288, 321-367) with plain integers."""
    cnt = np.ones(256, dtype=np.int64)
    lo, hi, pend, max_pend = 0, 0xFFFF, 0, 0
    out = np.empty(n, dtype=np.uint8)
    for i in range(n):
        total = 256 + i
        cum = np.concatenate(([0], np.cumsum(cnt)))
        rng = hi - lo + 1
        lows = lo + (cum[:-1] * rng) // total
        highs = lo + (cum[1:] * rng) // total - 1
        cand = np.nonzero((lows < 0x8000) & (highs >= 0x8000))[0]
        s = int(cand[0]) if cand.size else int(np.argmax(highs >= 0x8000 if lo < 0x8000 else highs >= lows))
        out[i] = s
        lo, hi = int(lows[s]), int(highs[s])
        cnt[s] += 1
        while True:
            if (lo ^ hi) & 0x8000 == 0:
                pend = 0
            elif (lo & 0x4000) and not (hi & 0x4000):
                pend += 1
                max_pend = max(max_pend, pend)
                lo &= 0x3FFF
                hi |= 0x4000
            else:
                break
            lo = (lo << 1) & 0xFFFF
            hi = ((hi << 1) | 1) & 0xFFFF
    return (out, max_pend) if report else out


GENERATORS = {"uniform": uniform, "and3": and3, "and2": and2, "mixed": mixed}


# ---------------------------------------------------------------- device side
def _mix_t(z):
    import torch

    def mul64(a, c):
        # torch has no uint64 arithmetic; int64 wraps the same way modulo 2^64
        return a * torch.tensor(c - (1 << 64) if c >= (1 << 63) else c, dtype=torch.int64, device=a.device)

    def lsr(a, k):
        # logical shift right on int64
        return (a >> k) & ((1 << (64 - k)) - 1)

    z = mul64(z ^ lsr(z, 30), _M1)
    z = mul64(z ^ lsr(z, 27), _M2)
    return z ^ lsr(z, 31)


def uniform_device(seed: int, n: int, start: int = 0, device="cuda"):
    """Same stream as :func:`uniform`, generated on ``device`` (start and n multiples of 8)."""
    import torch

    assert start % 8 == 0 and n % 8 == 0
    out = torch.empty(n, dtype=torch.uint8, device=device)
    words = out.view(torch.int64)
    step = 1 << 24
    g = _GOLDEN - (1 << 64)
    s = seed & _MASK
    s = s - (1 << 64) if s >= (1 << 63) else s
    for a in range(0, n // 8, step):
        b = min(n // 8, a + step)
        idx = torch.arange(start // 8 + a + 1, start // 8 + b + 1, dtype=torch.int64, device=device)
        words[a:b] = _mix_t(idx * g + s)
    return out


def and3_device(seed: int, n: int, start: int = 0, device="cuda"):
    return (uniform_device(seed, n, start, device) & uniform_device(seed + 1, n, start, device)
            & uniform_device(seed + 2, n, start, device))


def mixed_device(seed: int, n: int, start: int = 0, device="cuda"):
    """Same stream as :func:`mixed` (start and n multiples of 65536)."""
    import torch

    assert start % 65536 == 0 and n % 65536 == 0
    u = uniform_device(seed, n, start, device).view(-1, 65536)
    out = torch.empty_like(u)
    first = (start // 65536) % 4
    rows = torch.arange(u.shape[0], device=device)
    kind = (rows + first) % 4
    out[kind == 0] = u[kind == 0]
    del u
    a3 = and3_device(seed + 10, n, start, device).view(-1, 65536)
    out[kind == 1] = a3[kind == 1]
    del a3
    out[kind == 2] = 0
    a2 = (uniform_device(seed + 20, n, start, device) & uniform_device(seed + 21, n, start, device)).view(-1, 65536)
    out[kind == 3] = a2[kind == 3]
    del a2
    return out.view(-1)
