"""Per-kernel SASS evidence from the built library: cuobjdump -sass lib/libwwz_b200.so, instruction counts of the
mnemonics that prove what runs where (UBLKCP = 1-D TMA bulk copy, SYNCS = mbarrier, DMMA = FP64 tensor pipe,
DFMA/DMUL/DADD = FP64 CUDA-core pipe, MUFU = SFU).  Writes profiles/r02_sass_summary.txt."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pyleoclim_util_b200", "lib", "libwwz_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
kern, counts = None, collections.OrderedDict()
KEYS = ["UBLKCP", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "MUFU", "LDS", "STS", "ATOM", "RED", "SHFL", "total"]
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*\)$", "", kern).replace("wwz::", "")
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        op = m.group(1)
        counts[kern]["total"] += 1
        for k in KEYS[:-1]:
            if op.startswith(k):
                counts[kern][k] += 1
lines = ["# cuobjdump -sass pyleoclim_util_b200/lib/libwwz_b200.so  (tools/sass_summary.py); architectures: %s" % ", ".join(arch),
         "# UBLKCP = cp.async.bulk (1-D TMA), SYNCS = mbarrier ops, DMMA = FP64 tensor-pipe MMA (mma.sync m8n8k4 f64),",
         "# DFMA/DMUL/DADD = FP64 CUDA-core pipe.  Static instruction counts per kernel (not execution counts).",
         "%-64s " % "kernel" + " ".join("%7s" % k for k in KEYS)]
tot = collections.Counter()
for k, c in counts.items():
    lines.append("%-64s " % k[:64] + " ".join("%7d" % c[x] for x in KEYS))
    tot.update(c)
lines.append("%-64s " % "ALL KERNELS" + " ".join("%7d" % tot[x] for x in KEYS))
txt = "\n".join(lines) + "\n"
open(os.path.join(ROOT, "profiles", "r02_sass_summary.txt"), "w").write(txt)
print(txt)
