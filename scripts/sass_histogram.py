#!/usr/bin/env python
"""Opcode histogram per kernel from the built objects (cuobjdump -sass): the evidence that the sm_100a-specific instructions the
design relies on are really in the binaries (packed FP32x2 FFMA2/FMUL2/FADD2, bulk async copy UBLKCP + mbarrier SYNCS, VOTE, ...).
Usage: python scripts/sass_histogram.py [build/obj/*.o] > profiles/sass_rNN.txt"""
import collections
import glob
import re
import subprocess
import sys

WATCH = ["FFMA2", "FMUL2", "FADD2", "UBLKCP", "SYNCS", "MUFU", "F2I", "I2F", "I2FP", "LDG", "STG", "LDS", "STS", "ATOMS", "ATOMG", "RED", "VOTE", "SHFL", "BAR",
         "IMAD", "FFMA", "LDL", "STL", "LDC", "BRA"]


def main():
    objs = sys.argv[1:] or sorted(glob.glob("build/obj/*.o"))
    print("# SASS opcode histogram per kernel (static instruction counts; cuobjdump -sass of the shipped objects, sm_100a)")
    for o in objs:
        try:
            txt = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        except FileNotFoundError:
            print("cuobjdump not found"); return
        kern = None; hist = {}
        for ln in txt.splitlines():
            m = re.search(r"Function : (\S+)", ln)
            if m:
                kern = m.group(1); hist[kern] = collections.Counter(); continue
            m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
            if m and kern:
                hist[kern][m.group(1)] += 1
        if not hist:
            continue
        print(f"\n## {o}")
        for k, h in hist.items():
            short = subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()
            short = re.sub(r"\(anonymous namespace\)::|_GLOBAL__N__\w+::", "", short)
            short = re.sub(r"^void ", "", short)
            depth = 0; cut = len(short)
            for i, ch in enumerate(short):                      # drop the argument list, keep template arguments
                if ch == "<": depth += 1
                elif ch == ">": depth -= 1
                elif ch == "(" and depth == 0: cut = i; break
            short = short[:cut]
            tot = sum(h.values())
            keys = [w for w in WATCH if h.get(w)]
            print(f"{short:70s} total {tot:5d}  " + " ".join(f"{w}={h[w]}" for w in keys))


if __name__ == "__main__":
    main()
