"""Per-object counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (profiles/rNN_sass_summary.md).

    python tools/sass_summary.py > profiles/r02_sass_summary.md        # needs the built objects in km_unet_b200/_build
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "km_unet_b200", "_build")
MNEMONICS = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "REDUX", "MUFU.EX2", "STG.E.128", "LDG.E.128", "LDS.128", "HMMA", "ATOMG", "REDG"]


def main():
    print("# SASS evidence, round 2 build (`cuobjdump -sass km_unet_b200/_build/*.o`, sm_100a)\n")
    print("`UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld (TMEM -> registers), `UTCBAR` = tcgen05.commit, `UTMALDG` = cp.async.bulk.tensor (TMA "
          "tiled load), `UBLKCP` = cp.async.bulk (linear bulk copy), `SYNCS` = mbarrier ops, `REDUX` = redux.sync, `ATOMG` / `REDG` = global "
          "atomics (only DySample's dX scatter is left).  No `HMMA` (mma.sync) anywhere: every tensor-core instruction is tcgen05.\n")
    print("| object | kernels | " + " | ".join(MNEMONICS) + " |")
    print("|---|---|" + "---|" * len(MNEMONICS))
    for f in sorted(os.listdir(BUILD)):
        if not f.endswith(".o"):
            continue
        out = subprocess.run(["cuobjdump", "-sass", os.path.join(BUILD, f)], capture_output=True, text=True).stdout
        kernels = len(re.findall(r"^\s*Function : ", out, flags=re.M))
        cnt = collections.Counter()
        for line in out.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                for k in MNEMONICS:
                    if op.startswith(k):
                        cnt[k] += 1
        print(f"| {f} | {kernels} | " + " | ".join(str(cnt[k]) if cnt[k] else "" for k in MNEMONICS) + " |")


if __name__ == "__main__":
    sys.exit(main())
