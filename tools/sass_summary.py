#!/usr/bin/env python
"""Per-kernel SASS evidence for the Blackwell-native claims: counts of the tcgen05 / TMEM / TMA mnemonics in the shipped
lib/libmultb200.so (runs without a GPU: `cuobjdump -sass`).  Writes profiles/sass_summary.txt.

  UTCHMMA / UTCQMMA ...  tcgen05.mma (the 5th-generation tensor-core instruction; HMMA/QMMA = legacy mma.sync)
  LDTM / STTM            tcgen05.ld / tcgen05.st (tensor-memory accumulator access)
  UTMALDG / UTMASTG / UTMAREDG   cp.async.bulk.tensor load / store / reduce (TMA)
  LDGSTS                 cp.async (non-bulk) operand staging
  SYNCS                  mbarrier operations
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "multimodal-transformer-robustness_b200", "lib", "libmultb200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTMAPF", "LDGSTS",
       "SYNCS", "HMMA", "QMMA", "IMMA", "FFMA", "MUFU", "BAR.SYNC", "ACQBULK", "UBLKCP"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    fn = None
    counts = collections.OrderedDict()
    arch = set(re.findall(r"arch = (sm_\w+)", out))
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            fn = m.group(1)
            counts[fn] = collections.Counter()
            continue
        if fn is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1)
            counts[fn]["_total"] += 1
            for p in PAT:
                if op.startswith(p):
                    counts[fn][p] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    lines = [f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (arch {', '.join(sorted(arch))}); tools/sass_summary.py",
             "# tcgen05.mma = UTC*MMA, tcgen05.ld/st = LDTM/STTM, TMA = UTMALDG/UTMASTG/UTMAREDG, cp.async = LDGSTS, mbarrier = SYNCS", ""]
    for name, (raw, c) in zip(demangle, counts.items()):
        short = re.sub(r"\(.*", "", name)
        used = [f"{p}={c[p]}" for p in PAT if c[p]]
        lines.append(f"{short:60s} instr={c['_total']:6d}  " + " ".join(used))
    txt = "\n".join(lines) + "\n"
    with open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w") as f:
        f.write(txt)
    sys.stdout.write(txt)


if __name__ == "__main__":
    main()
