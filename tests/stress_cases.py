"""Random inputs for the randomised round-trip checks (tests/test_gpu_stress.py, tools/gpu_stress.py)."""
import numpy as np

import vectors as T

EDGE_SIZES = [0, 1, 2, 3, 100, 4095, 4096, 4097, 32767, 32768, 32769, 65536, 131071, 131072, 131073]


def make(rng, n: int, raw: bytes | None = None) -> bytes:
    """n bytes of one of seven kinds: random, small alphabet, a repeated unit, the reference's fixture, zeros, a
    patchwork of those with runs of zeros, LCG text."""
    raw = raw if raw is not None else T.fixture_raw()
    kind = int(rng.integers(0, 7))
    if kind == 0:
        return rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
    if kind == 1:
        return rng.integers(97, 97 + int(rng.integers(2, 30)), size=n, dtype=np.uint8).tobytes()
    if kind == 2:
        unit = rng.integers(0, 256, size=int(rng.integers(1, 70000)), dtype=np.uint8).tobytes()
        return (unit * (n // len(unit) + 1))[:n]
    if kind == 3:
        o = int(rng.integers(0, max(1, len(raw) - n)))
        return (raw * (n // len(raw) + 2))[o:o + n]
    if kind == 4:
        return bytes(n)
    if kind == 5:
        parts, left = [], n
        while left > 0:
            m = min(left, int(rng.integers(1, 50000)))
            parts.append(make(rng, m, raw) if rng.integers(0, 4) else bytes(m))
            left -= m
        return b"".join(parts)[:n]
    return T.gen("G5", n)


def size(rng, maxlog: int = 22) -> int:
    if rng.integers(0, 3) == 0:
        return int(rng.choice(EDGE_SIZES))
    return int(rng.integers(0, 1 << int(rng.integers(4, maxlog))))
