"""SASS / resource evidence of the built library, readable without a GPU:
per kernel the register count, shared memory, spills (cuobjdump --dump-resource-usage) and the count of the mnemonics
that show which hardware paths the code uses (bulk TMA copies, cp.async, mbarrier, packed fp32, SFU, ballots, ...).
usage: python tools/sass_evidence.py > profiles/<tag>_sass_evidence.md"""
import collections, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "vk_gaussian_splatting_b200" / "lib" / "libvkgs_b200.so"
KEYS = [("UBLKCP", "cp.async.bulk (TMA bulk copy global->shared)"), ("SYNCS", "mbarrier arrive / try_wait"), ("LDGSTS", "cp.async (per-thread async gather)"),
        ("FFMA2", "packed fp32 FMA (two results per issue)"), ("FMUL2", "packed fp32 multiply"), ("FADD2", "packed fp32 add"), ("FFMA", "scalar FFMA (incl. FFMA2)"),
        ("MUFU.EX2", "SFU exp2"), ("MUFU.RCP", "SFU reciprocal"), ("MUFU.RSQ", "SFU rsqrt"), ("VOTE", "warp ballot"), ("MATCH", "match.any"),
        ("REDUX", "warp reduce"), ("ATOMS", "shared-memory atomics"), ("ATOMG", "global atomics"), ("RED.", "global reductions"),
        ("STL", "local-memory stores (spills)"), ("LDL", "local-memory loads (spills)"), ("BAR.SYNC", "CTA barriers"), ("LDS", "shared loads"), ("STS", "shared stores"),
        ("LDG", "global loads"), ("STG", "global stores")]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except Exception:
        return name


def short(name):
    d = demangle(name)
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"vkgs::", "", d)
    d = re.sub(r"^void ", "", d)
    return re.sub(r"\(.*\)$", "", d)


def main():
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", str(LIB)], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = dict(re.findall(r"(\w+):(\d+)", line))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True).stdout
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if cur and m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for k, _ in KEYS:
                if op.startswith(k):
                    counts[cur][k] += 1
    print("# SASS / resource evidence of `libvkgs_b200.so` (sm_100a), produced on the build container by `tools/sass_evidence.py`\n")
    print("Mnemonics: " + "; ".join(f"`{k}` = {d}" for k, d in KEYS) + ".\n")
    print("| kernel | regs | smem (static) | spill st/ld | instr | " + " | ".join(k for k, _ in KEYS) + " |")
    print("|---|---|---|---|---|" + "---|" * len(KEYS))
    for fn, c in counts.items():
        if "k_" not in fn:
            continue
        u = usage.get(fn, {})
        print(f"| `{short(fn)}` | {u.get('REG', '?')} | {u.get('SHARED', '?')} | {c['STL']}/{c['LDL']} | {c['_total']} | "
              + " | ".join(str(c[k]) if c[k] else "" for k, _ in KEYS) + " |")


if __name__ == "__main__":
    sys.exit(main())
