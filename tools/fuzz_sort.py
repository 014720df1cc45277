"""Exploratory fuzz of the stand-alone radix sort: sizes around the partition boundaries, adversarial key distributions;
vkgs_sort_pairs against numpy's stable argsort.  python tools/fuzz_sort.py [trials]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import vk_gaussian_splatting_b200 as g


def run(trials, r, log=print):
    bad = 0
    rng = np.random.default_rng(4242)
    base = [0, 1, 2, 31, 32, 33, 255, 256, 257, 4095, 4096, 4097, 8191, 8192, 8193, 12288, 65535, 65536, 65537, 1 << 20, (1 << 20) + 1, 3_000_001]
    for t in range(trials):
        n = int(rng.choice(base)) if rng.random() < 0.6 else int(rng.integers(0, 600_000))
        kind = int(rng.integers(0, 9))
        if kind == 0:
            k = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
        elif kind == 1:
            k = np.full(n, rng.integers(0, 1 << 32), np.uint32)
        elif kind == 2:
            k = rng.integers(0, 2, n).astype(np.uint32) * np.uint32(0x80000000)
        elif kind == 3:
            k = np.sort(rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))
        elif kind == 4:
            k = np.sort(rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))[::-1].copy()
        elif kind == 5:
            k = (rng.integers(0, 1 << int(rng.integers(1, 33)), n, dtype=np.uint64)).astype(np.uint32)      # few low bits
        elif kind == 6:
            k = (rng.integers(0, 256, n, dtype=np.uint64) << int(rng.choice([0, 8, 16, 24]))).astype(np.uint32)  # one live byte
        elif kind == 7:
            k = np.where(rng.random(n) < 0.999, 0x12345678, rng.integers(0, 1 << 32, n, dtype=np.uint64)).astype(np.uint32)  # one hot bin
        else:
            k = np.arange(n, dtype=np.uint32) * np.uint32(2654435761)
        v = rng.permutation(n).astype(np.uint32)
        ks, vs, _ = r.sort_pairs(k, v)
        order = np.argsort(k, kind="stable")
        if not (np.array_equal(ks, k[order]) and np.array_equal(vs, v[order])):
            log(f"trial {t}: n={n} kind={kind} -> MISMATCH", flush=True)
            bad += 1
    return bad


if __name__ == "__main__":
    n_trials = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    print(f"mismatching {run(n_trials, g.GaussianSplatting(0))} of {n_trials}")
