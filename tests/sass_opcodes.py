"""Developer tool: per-kernel counts of the Blackwell-specific SASS opcodes in the built library
(tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, tcgen05.commit -> UTCBAR).
    python tests/sass_opcodes.py > profiles/sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "lidar_layout_b200", "liblidm_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTMAPF", "HMMA", "SYNCS", "MUFU.EX2", "MUFU.TANH", "FFMA2", "ACQBULK"]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
dem = {}
counts = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                counts[cur][o] += 1
names = list(counts)
d = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines() if names else []
print(f"# {os.path.relpath(lib, ROOT)}: {len(names)} kernels; cuobjdump -sass | per-kernel opcode counts (sm_100a)")
print("# " + " ".join(f"{o:>8s}" for o in ["instrs"] + OPS) + "  kernel")
tot = collections.Counter()
for n, dn in zip(names, d or names):
    c = counts[n]
    tot.update(c)
    if not any(c[o] for o in OPS[:9]):
        continue
    short = dn.split(">(")[0] + (">" if ">(" in dn else "")
    short = re.sub(r"\((int|bool)\)", "", short).replace("lidm::<unnamed>::", "").replace("void ", "")
    print("  " + " ".join(f"{c[o]:8d}" for o in ["_total"] + OPS) + "  " + short)
print("  " + " ".join(f"{tot[o]:8d}" for o in ["_total"] + OPS) + "  TOTAL (all kernels, including those without tensor/TMA opcodes)")
