"""Golden vectors (tests/golden/, produced from the reference by make_golden.py)."""
from __future__ import annotations

import hashlib
import json
import os

import numpy as np

from gpuar_b200 import datagen as D

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VECTORS = json.load(open(os.path.join(GOLDEN, "vectors.json")))
SWEEP = json.load(open(os.path.join(GOLDEN, "sweep.json")))

SMALL = [k for k, v in VECTORS.items() if v["n"] <= (1 << 20) + 8192]   # seconds on one CPU core
LARGE = [k for k in VECTORS if k not in SMALL]


def md5(a) -> str:
    return hashlib.md5(np.ascontiguousarray(a).tobytes()).hexdigest()


def make_input(rec: dict) -> np.ndarray:
    g = rec["gen"]
    if g == "bytes":
        return np.frombuffer(bytes.fromhex(rec["hex"]), dtype=np.uint8).copy()
    if g == "file":
        return np.fromfile(os.path.join(GOLDEN, rec["file"]), dtype=np.uint8)
    if g == "zeros":
        return D.zeros(rec["n"])
    if g == "round_robin":
        return D.round_robin(rec["n"])
    if g == "adversarial_x4":
        return np.tile(D.adversarial_packet(), 4)
    return D.GENERATORS[g](rec["seed"], rec["n"])


def stored_payload(name: str):
    p = os.path.join(GOLDEN, f"{name}.payload")
    return np.fromfile(p, dtype=np.uint8) if os.path.exists(p) else None
