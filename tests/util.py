import hashlib
import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def seeded_radiosity(P, seed):
    """Same generator as tests/golden/make_golden.py (integer hash, portable)."""
    x = np.arange(P * 3, dtype=np.uint64) * np.uint64(6364136223846793005) + np.uint64(1442695040888963407 + seed)
    x ^= x >> np.uint64(29)
    x = (x * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x ^= x >> np.uint64(32)
    q = (x % np.uint64(9)).astype(np.float32) / np.float32(4.0)
    zero = ((x >> np.uint64(8)) % np.uint64(3)) == 0
    q[zero] = 0
    return q.reshape(P, 3)


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / d) if d > 0 else float(np.linalg.norm(a - b))
