#!/usr/bin/env python3
"""Static evidence of what the built library contains (no GPU needed): per kernel, the SASS instruction count with the
tcgen05 / TMA / TMEM mnemonics of /opt/skills/guides/B200_PROFILING.md (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2,
UTMALDG = TMA tensor load, LDTM = tcgen05.ld) from ``cuobjdump -sass``, and registers / stack / spills from a
``-Xptxas -v`` compile of the same sources into a scratch file.

    python tools/sass_summary.py > profiles/<round>_sass_summary.md
"""
from __future__ import annotations

import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from tokenize_audio_b200 import build as B  # noqa: E402

MNEMONICS = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "MUFU", "HMMA", "FFMA")


def demangle(name: str) -> str:
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    return re.sub(r"\(.*", "", out).replace("void ", "")


def sass_counts(lib: str):
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    fn, cnt = None, collections.defaultdict(collections.Counter)
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = demangle(m.group(1))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if fn is None or not m:
            continue
        op = m.group(1)
        cnt[fn]["all"] += 1
        base = op.split(".")[0]
        if base in MNEMONICS:
            cnt[fn][base + (".2CTA" if "2CTA" in op else "")] += 1
    return cnt


def ptxas_info():
    with tempfile.TemporaryDirectory() as d:
        cmd = ["nvcc", *B.NVCC_FLAGS, "-Xptxas=-v", "-o", os.path.join(d, "scratch.so"), os.path.join(B.CSRC, "mimi_b200.cu")]
        err = subprocess.run(cmd, capture_output=True, text=True, check=True).stderr
    info = {}
    for blk in re.split(r"ptxas info\s+: Compiling entry function '", err)[1:]:
        name = demangle(blk.split("'")[0])
        regs = re.search(r"Used (\d+) registers", blk)
        spill = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", blk)
        stack = re.search(r"(\d+) bytes stack frame", blk)
        info[name] = (int(regs.group(1)) if regs else -1, int(stack.group(1)) if stack else 0,
                      int(spill.group(1)) if spill else 0, int(spill.group(2)) if spill else 0)
    return info


def main():
    lib = B.build()
    cnt, info = sass_counts(lib), ptxas_info()
    tot = collections.Counter()
    for c in cnt.values():
        tot.update(c)
    print("# SASS / ptxas summary of libmimi_b200.so (sm_100a)\n")
    print("`python tools/sass_summary.py`: `cuobjdump -sass` mnemonic counts per kernel (static, not executed counts) and "
          "`-Xptxas -v` resource lines.\n")
    print("Library totals: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items()) if k not in ("all", "FFMA", "MUFU")) +
          f"; {len(cnt)} kernels, {tot['all']} instructions.\n")
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMALDG.2CTA", "LDTM", "UTCBAR", "UTCBAR.2CTA", "FFMA", "MUFU"]
    print("| kernel | SASS instr. | " + " | ".join(cols) + " | regs | stack B | spill st / ld B |")
    print("|---|---:|" + "---:|" * len(cols) + "---:|---:|---:|")
    for fn, c in sorted(cnt.items(), key=lambda kv: -kv[1]["all"]):
        r = info.get(fn, (-1, 0, 0, 0))
        print(f"| `{fn}` | {c['all']} | " + " | ".join(str(c[k]) if c[k] else "" for k in cols) + f" | {r[0]} | {r[1]} | {r[2]} / {r[3]} |")


if __name__ == "__main__":
    main()
