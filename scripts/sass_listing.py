"""Per-kernel SASS summary of the in-tree extension (no GPU needed): instruction count, registers / shared memory,
the mnemonics that prove which hardware path a kernel uses, and the 12 most frequent mnemonics.

    python scripts/sass_listing.py > profiles/sass_listing_r1.txt

Key mnemonics (sm_100a): UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = TMEM
alloc/dealloc, UTMALDG = TMA tensor load (cp.async.bulk.tensor), UTMAPF = tensormap prefetch, SYNCS = mbarrier ops,
LDGMC / STGMC(REDGMC) = multimem.ld_reduce / multimem.st (NVSwitch multicast), REDG = red.global (vector atomics),
ATOMS / ATOMG = shared / global atomics, MUFU = special function unit, SHFL = warp shuffles, BAR = CTA barriers.
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(REPO, "mine_b200", "ops", "_mine_b200_cuda.so")
KEY = ["UTCHMMA", "LDTM", "UTCBAR", "UTCATOMSWS", "UTMALDG", "UTMAPF", "SYNCS", "LDGMC", "STGMC", "REDGMC", "REDG", "ATOMS",
       "ATOMG", "RED", "MUFU", "SHFL", "BAR", "LDS", "STS", "LDG", "STG", "HMMA", "FFMA", "MEMBAR"]


def demangle(names):
    out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.strip().split("\n")
    return dict(zip(names, out))


def main():
    cuobjdump = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "bin", "cuobjdump")
    sass = subprocess.run([cuobjdump, "-sass", SO], capture_output=True, text=True).stdout
    res = subprocess.run([cuobjdump, "-res-usage", SO], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = " ".join(re.findall(r"(?:REG|STACK|SHARED|LOCAL):\d+", line))
            cur = None
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    names = demangle(list(kernels))
    print(__doc__.split("Key mnemonics")[0].strip().split("\n")[0])
    print("source: cuobjdump -sass / -res-usage %s\n" % os.path.relpath(SO, REPO))
    for k, cnt in sorted(kernels.items(), key=lambda kv: names[kv[0]]):
        short = re.sub(r"\(.*", "", names[k]).replace("mine::", "").replace("(anonymous namespace)::", "")
        total = sum(cnt.values())
        keys = ", ".join("%s x%d" % (m, cnt[m]) for m in KEY if cnt.get(m))
        top = ", ".join("%s %d" % (m, n) for m, n in cnt.most_common(12))
        print("%s\n    %d instructions; %s\n    key: %s\n    top: %s" % (short, total, usage.get(k, ""), keys, top))


if __name__ == "__main__":
    main()
