import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def bits_to_f32(a):
    return np.asarray(a, dtype=np.uint32).view(np.float32)


def f32_bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def load_glm_golden():
    with open(GOLDEN / "glm_golden.json") as f:
        return json.load(f)


def ulp_diff(a, b):
    """Distance in units of last place between two float32 arrays of the same sign."""
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)
