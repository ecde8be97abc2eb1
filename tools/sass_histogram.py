"""Opcode histograms of the built kernels (cuobjdump -sass on whvi_b200/_obj/*.o): the committed evidence that the hot
kernels are Blackwell-native -- UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTCBAR, UBLKCP / UTMALDG (bulk/TMA
copies), SYNCS (mbarrier), FADD2/FMUL2/FFMA2 (packed fp32x2).    python tools/sass_histogram.py > profiles/r02_sass.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FADD2", "FMUL2", "FFMA2",
         "FADD", "FMUL", "FFMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "HMMA", "LDL", "STL"]


def main():
    objs = sorted((ROOT / "whvi_b200" / "_obj").glob("*.o"))
    if not objs:
        sys.exit("build first: python -m whvi_b200.build")
    print("# cuobjdump -sass opcode counts per kernel (static instruction counts; sm_100a objects of libwhvi_b200.so)")
    print("# columns: " + " ".join(WATCH) + " | total")
    for obj in objs:
        sass = subprocess.run(["cuobjdump", "-sass", str(obj)], capture_output=True, text=True).stdout
        kernels, cur = collections.OrderedDict(), None
        for line in sass.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = kernels.setdefault(m.group(1), collections.Counter())
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m and cur is not None:
                cur[m.group(1)] += 1
        print(f"\n== {obj.name}")
        for name, c in kernels.items():
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            if dem.endswith(")"):   # drop the parameter list (the last top-level parenthesised group), keep template arguments
                depth = 0
                for i in range(len(dem) - 1, -1, -1):
                    depth += dem[i] == ")"
                    depth -= dem[i] == "("
                    if depth == 0:
                        dem = dem[:i]
                        break
            dem = dem.replace("whvi::", "").replace("void ", "").replace("(int)", "").replace("(bool)", "")
            cols = " ".join(f"{c.get(w, 0)}" for w in WATCH)
            print(f"{dem[:100]:100s} {cols} | {sum(c.values())}")


if __name__ == "__main__":
    main()
