"""Count the Blackwell-specific SASS mnemonics per kernel of libstlpose_b200.so (B200_PROFILING.md: "What proves a
Blackwell-native kernel").  `python tools/sass_census.py > profiles/r02_sass_census.txt` - runs here, no GPU needed."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "stlpose_b200", "libstlpose_b200.so")
PATTERNS = collections.OrderedDict([
    ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
    ("UTCBAR", r"\bUTCBAR"), ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"),
    ("SYNCS", r"\bSYNCS"), ("HMMA", r"\bHMMA"), ("HGMMA", r"\bHGMMA"), ("ATOM/RED.F32", r"\b(ATOMG|ATOM|RED|REDG)\.[A-Z.]*F32"),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        for name, pat in PATTERNS.items():
            if re.search(pat, line):
                counts[cur][name] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(order), capture_output=True, text=True).stdout.splitlines()
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (sm_100a)")
    print("# UTCHMMA = tcgen05.mma (kind::f16), .2CTA = cta_group::2; LDTM/STTM = tcgen05.ld/st; UTCBAR = tcgen05.commit;")
    print("# UTMALDG/UTMASTG = TMA tensor load/store; SYNCS = mbarrier ops; HMMA = legacy mma.sync (must be 0)")
    hdr = f"{'kernel':78s} " + " ".join(f"{n:>12s}" for n in PATTERNS)
    print(hdr)
    total = collections.Counter()
    for fn, name in zip(order, demangle):
        c = counts[fn]
        total.update(c)
        short = re.sub(r"\(anonymous namespace\)::", "", name)
        short = re.sub(r"\(.*", "", short).replace("void ", "")
        print(f"{short[:78]:78s} " + " ".join(f"{c.get(n, 0):12d}" for n in PATTERNS))
    print(f"{'TOTAL (' + str(len(order)) + ' kernels)':78s} " + " ".join(f"{total.get(n, 0):12d}" for n in PATTERNS))


if __name__ == "__main__":
    main()
