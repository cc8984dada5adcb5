"""SASS mnemonic summary of libgphm.so for profiles/ (no GPU needed: `cuobjdump -sass` reads the in-tree library).

    python tools/sass_summary.py [profiles/r02_sass_summary.txt]

One line per kernel (template instantiations summed): the FP64 / shared-memory / barrier mnemonics that the roofline
discussion in DESIGN.md refers to and every Blackwell-native one (tcgen05 = UTC*MMA / LDTM / UTCBAR / UTCATOMSWS, TMA = UTMALDG,
mbarrier = SYNCS), plus system-scope fences / atomics of the NVLink peer exchange and the cluster barrier of the Schur recursion.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussian-process-slover-for-high-freq-pde_b200", "libgphm.so")
KEEP = re.compile(r"^(DFMA|DMUL|DADD|DMMA|LDS|STS|BAR|MUFU\.RCP64H|LDGSTS|UTC|LDTM|STTM|UTMA|SYNCS|MEMBAR|ATOMG|ATOMS|RED|UCGABAR|CGABAR|"
                  r"ST\.E.*SYS|LD\.E.*SYS|NANOSLEEP|ERRBAR|FENCE)")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], check=True, capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    inst = collections.Counter()
    cur = None
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            dem = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            dem = re.sub(r"(\(anonymous namespace\)|<unnamed>)::", "", dem)
            dem = re.sub(r"^void\s+", "", dem)
            name = re.sub(r"[<(].*", "", dem).split("::")[-1]
            cur = per.setdefault(name, collections.Counter())
            inst[name] += 1
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            if KEEP.match(op):
                # fold operand-size / cache-hint suffixes that do not matter here
                op = re.sub(r"^(LDS|STS)\..*", r"\1", op)
                op = re.sub(r"^BAR\.(SYNC|ARV|RED).*", r"BAR.\1", op)
                cur[op] += 1
    total = collections.Counter()
    with open(out, "w") as f:
        f.write("# SASS mnemonics of libgphm.so (cuobjdump -sass, sm_100a), summed over the template instantiations of each kernel\n"
                "# (tools/sass_summary.py).  Blackwell-native: UTCIMMA = tcgen05.mma kind::i8, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA),\n"
                "# UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc, SYNCS = mbarrier; DMMA = mma.sync f64 (the native-FP64 GEMM);\n"
                "# LDGSTS = cp.async; ST / LD ... .SYS and MEMBAR.*.SYS = system-scope release / acquire (flags of the NVLink peer exchange);\n"
                "# UCGABAR = cluster barrier, MEMBAR.*.GPU = the device-scope fences of the Schur hand-over.\n\n")
        for name, c in per.items():
            f.write("%-34s x%-3d %s\n" % (name, inst[name], "  ".join("%s=%d" % kv for kv in c.most_common())))
            total.update(c)
        f.write("\nTOTAL  %s\n" % "  ".join("%s=%d" % kv for kv in total.most_common()))
    print(open(out).read()[-1500:])


if __name__ == "__main__":
    main()
