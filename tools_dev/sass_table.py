"""Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library (tcgen05 MMA = UTC*MMA,
TMEM loads / stores = LDTM / STTM, TMA tensor loads = UTMALDG, bulk copies / prefetches = UBLKCP / UBLKPF,
tcgen05.commit = UTCBAR, setmaxnreg = USETMAXREG), plus an excerpt of one kernel.
usage: python tools_dev/sass_table.py [lib.so] [kernel substring for the excerpt]"""
import os, re, subprocess, sys
from collections import Counter, OrderedDict
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "deep_cartograph_b200", "lib", "libdcg_b200.so")
pick = sys.argv[2] if len(sys.argv) > 2 else "cov_i8_fused_kernelILb1"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
names = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "USETMAXREG", "UCGABAR"]
rx = re.compile(r"\b(" + "|".join(names) + r")[A-Z_]*")
tab, fn, excerpt = OrderedDict(), None, []
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); tab[fn] = Counter(); continue
    m = rx.search(line)
    if m and fn:
        tab[fn][m.group(1)] += 1
        if pick in fn and len(excerpt) < 60:
            excerpt.append(re.sub(r"/\*[0-9a-f]{16}\*/", "", re.sub(r"^\s*/\*([0-9a-f]{4,5})\*/", r"\1", line)).rstrip())
dem = subprocess.run(["c++filt"], input="\n".join(tab), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.relpath(so, root)} (sm_100a), cuobjdump -sass; kernels without any of these mnemonics are omitted")
print("# " + " ".join(f"{n:>10s}" for n in names) + "  kernel")
for (fn, c), d in zip(tab.items(), dem):
    if sum(c.values()) == 0:
        continue
    d = re.sub(r"dcg::\(anonymous namespace\)::|\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d)
    print("  " + " ".join(f"{c[n]:10d}" for n in names) + "  " + d)
print(f"\n# excerpt ({pick}): the Blackwell-specific instructions in program order")
print("\n".join(excerpt))
