"""Mutation fuzz of the scene loaders (csrc/host_loader.cpp): random byte flips, truncations, header number edits and
appended garbage on the fixture files; every mutant is loaded in a child process so a crash is seen as a signal, and a
mutant that loads must yield arrays of consistent sizes. No GPU needed.   python tools/fuzz_loader.py [mutants per file] [seed]"""
import os, subprocess, sys, tempfile, re
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np

CHILD = r'''
import sys
sys.path.insert(0, %r)
import numpy as np
import vk_gaussian_splatting_b200 as g
ok = err = 0
for p in sys.argv[1:]:
    try:
        s = g.load_scene(p)
        n = s.size()
        assert s.positions.shape == (n, 3) and s.scale.shape == (n, 3) and s.rotation.shape == (n, 4) and s.opacity.shape == (n,)
        assert s.f_dc.shape == (n, 3) and s.f_rest.shape[0] == n and s.f_rest.shape[1] in (0, 9, 24, 45)
        float(np.nansum(s.positions)) + float(np.nansum(s.f_rest))   # touch every page
        ok += 1
    except g.VkgsError:
        err += 1
print("loaded", ok, "rejected", err)
''' % str(ROOT)


def mutate(data: bytes, rng) -> bytes:
    b = bytearray(data)
    kind = rng.integers(0, 6)
    if kind == 0 and len(b) > 1:      # truncate
        return bytes(b[:rng.integers(0, len(b))])
    if kind == 1:                      # byte flips anywhere
        for _ in range(int(rng.integers(1, 20))):
            b[rng.integers(0, len(b))] = rng.integers(0, 256)
    if kind == 2:                      # byte flips in the first 512 bytes (headers)
        for _ in range(int(rng.integers(1, 8))):
            b[rng.integers(0, min(512, len(b)))] = rng.integers(0, 256)
    if kind == 3:                      # edit a decimal number in the header (element counts, list lengths)
        head = bytes(b[:2048])
        nums = list(re.finditer(rb"\d+", head))
        if nums:
            m = nums[rng.integers(0, len(nums))]
            new = str(int(rng.choice([0, 1, 2**31 - 1, 2**31, 2**32 - 1, 2**32, 2**63 - 1, 2**64 - 1, 10**30, int(m.group()) * 1000 + 7]))).encode()
            b[m.start():m.end()] = new
    if kind == 4:                      # append garbage
        b += bytes(rng.integers(0, 256, int(rng.integers(1, 4096)), dtype=np.uint8))
    if kind == 5 and len(b) > 64:     # overwrite a block with 0xff / 0x00
        i = rng.integers(0, len(b) - 32)
        b[i:i + 32] = bytes([int(rng.choice([0, 255]))]) * 32
    return bytes(b)


def main():
    per_file = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rng = np.random.default_rng(seed)
    files = sorted(p for p in (ROOT / "tests" / "golden").iterdir() if p.suffix in (".ply", ".spz", ".splat"))
    crashes = 0
    with tempfile.TemporaryDirectory() as td:
        for f in files:
            data = f.read_bytes()
            paths = []
            for i in range(per_file):
                p = Path(td) / f"{f.stem}_{i}{f.suffix}"
                p.write_bytes(mutate(data, rng))
                paths.append(str(p))
            for k in range(0, len(paths), 50):
                batch = paths[k:k + 50]
                pr = subprocess.run([sys.executable, "-c", CHILD, *batch], capture_output=True, text=True, timeout=600)
                if pr.returncode != 0:
                    # find the culprit one by one
                    for p in batch:
                        q = subprocess.run([sys.executable, "-c", CHILD, p], capture_output=True, text=True, timeout=120)
                        if q.returncode != 0:
                            crashes += 1
                            keep = ROOT / "gpurun_out" / ("crash_" + Path(p).name)
                            keep.parent.mkdir(exist_ok=True)
                            keep.write_bytes(Path(p).read_bytes())
                            print("CRASH", f.name, "rc", q.returncode, "kept as", keep, (q.stderr or "")[-300:].replace("\n", " | "))
                else:
                    print(f.name, pr.stdout.strip())
            for p in paths:
                os.unlink(p)
    print("crashing mutants:", crashes)
    return 1 if crashes else 0


if __name__ == "__main__":
    sys.exit(main())
