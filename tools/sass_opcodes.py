"""SASS opcode histogram per kernel of libdkg_b200.so (run here, no GPU): profiles/*_sass_opcodes.txt."""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "decoupled-kg_b200", "lib", "libdkg_b200.so")
COLS = ["UTCIMMA", "LDTM", "UBLKCP", "UTCBAR", "SYNCS", "CREDUX", "REDUX", "DMMA", "DFMA", "DADD", "DMUL", "FFMA2", "FFMA", "FSETP",
        "F2F", "LDG", "STG", "LDS", "STS", "SHFL", "ATOM", "ATOMS", "RED", "BAR", "MUFU", "LDGSTS", "ELECT"]


def main(out):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    hist, cur, order, k = {}, None, [], 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", names[k].replace("(anonymous namespace)::", "")); k += 1
            hist[cur] = collections.Counter(); order.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            hist[cur]["total"] += 1
            hist[cur][m.group(1)] += 1
    with open(out, "w") as f:
        f.write("# SASS opcode histogram of decoupled-kg_b200/lib/libdkg_b200.so (cuobjdump -sass, sm_100a), one row per kernel\n")
        f.write("# tcgen05.mma.kind::i8 -> UTCIMMA, tcgen05.ld -> LDTM, cp.async.bulk -> UBLKCP, tcgen05.commit / mbarrier -> UTCBAR / SYNCS,\n")
        f.write("# redux.sync.{min,max}.f32 -> CREDUX, elect.sync -> ELECT, fp64 mma.sync -> DMMA, packed fp32 fma -> FFMA2, cp.async -> LDGSTS.\n")
        f.write("# Regenerate: python tools/sass_opcodes.py profiles/<name>.txt\n")
        f.write("kernel | total | " + " | ".join(COLS) + "\n")
        for name in order:
            h = hist[name]
            f.write(f"{name} | {h['total']} | " + " | ".join(str(h[c]) for c in COLS) + "\n")


if __name__ == "__main__":
    main(sys.argv[1])
