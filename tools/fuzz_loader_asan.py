"""The loader mutation fuzz of tools/fuzz_loader.py under AddressSanitizer + UndefinedBehaviorSanitizer: csrc/host_loader.cpp is
compiled alone (it needs no CUDA) with tools/loader_asan_driver.cpp into tools/bin/loader_asan, which loads every mutant and
touches every array the scene view exposes. No GPU needed.   python tools/fuzz_loader_asan.py [mutants per file] [seed]"""
import importlib.util, os, subprocess, sys, tempfile
from pathlib import Path
import numpy as np

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "tools" / "bin" / "loader_asan"


def build():
    BIN.parent.mkdir(exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", f"-I{ROOT / 'include'}",
                    str(ROOT / "vk_gaussian_splatting_b200" / "csrc" / "host_loader.cpp"), str(ROOT / "tools" / "loader_asan_driver.cpp"),
                    "-lz", "-lpthread", "-o", str(BIN)], check=True)


def flagged(pr):
    return pr.returncode != 0 or "runtime error" in pr.stderr or "ERROR: AddressSanitizer" in pr.stderr


def main():
    per_file = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    build()
    spec = importlib.util.spec_from_file_location("fuzz_loader", ROOT / "tools" / "fuzz_loader.py")
    fl = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fl)
    rng = np.random.default_rng(seed)
    env = dict(os.environ, ASAN_OPTIONS="detect_leaks=0", UBSAN_OPTIONS="print_stacktrace=1")
    files = sorted(p for p in (ROOT / "tests" / "golden").iterdir() if p.suffix in (".ply", ".spz", ".splat"))
    reports = 0
    with tempfile.TemporaryDirectory() as td:
        for f in files:
            data = f.read_bytes()
            paths = []
            for i in range(per_file):
                p = Path(td) / f"{f.stem}_{i}{f.suffix}"
                p.write_bytes(fl.mutate(data, rng))
                paths.append(str(p))
            for k in range(0, len(paths), 100):
                batch = paths[k:k + 100]
                if flagged(subprocess.run([str(BIN), *batch], capture_output=True, text=True, timeout=900, env=env)):
                    for p in batch:
                        q = subprocess.run([str(BIN), p], capture_output=True, text=True, timeout=120, env=env)
                        if flagged(q):
                            reports += 1
                            keep = ROOT / "gpurun_out" / ("asan_" + Path(p).name)
                            keep.parent.mkdir(exist_ok=True)
                            keep.write_bytes(Path(p).read_bytes())
                            print("REPORT", f.name, "kept as", keep, "|", " ".join(l.strip() for l in q.stderr.splitlines() if "runtime error" in l or "ERROR" in l)[:600], flush=True)
            for p in paths:
                os.unlink(p)
    print("sanitizer reports:", reports)
    return 1 if reports else 0


if __name__ == "__main__":
    sys.exit(main())
