#!/usr/bin/env python
"""Static facts about every kernel of the library, no GPU needed: registers / spills / static shared memory from `ptxas -v`, instruction
counts and the mnemonics that show which hardware paths a kernel uses from `cuobjdump -sass` (tcgen05: UTCHMMA / UTCBAR / LDTM, TMA:
UTMALDG / UTMAPF, cp.async: LDGSTS, special-function unit: MUFU, NVLink multimem: MULTIMEM / LDGMC / STGMC ... whatever the file has).

    python tools/static_report.py > profiles/r1_static_kernels.md
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "wav2letter_pytorch_b200", "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
WATCH = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UTMAPF", "SYNCS", "LDGSTS", "MUFU", "SHFL", "ATOMS", "ATOMG", "RED", "MULTIMEM", "BAR"]


def demangle(names):
    if not names:                                             # (c++filt without arguments would wait on stdin)
        return {}
    out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout.splitlines()
    return dict(zip(names, (re.sub(r"\(.*", "", o).replace("void ", "").replace("w2l::", "") for o in out)))


def main():
    rows = []
    with tempfile.TemporaryDirectory() as tmp:
        for src in sorted(f for f in os.listdir(CSRC) if f.endswith(".cu")):
            obj = os.path.join(tmp, src.replace(".cu", ".o"))
            r = subprocess.run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                                "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj], capture_output=True, text=True)
            if r.returncode:
                sys.exit(r.stderr)
            info, cur = {}, None
            for line in r.stderr.splitlines():
                m = re.search(r"Compiling entry function '(\S+)'", line)
                if m:
                    cur = m.group(1)
                    info[cur] = {"regs": "?", "spill": 0, "smem": 0}
                m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
                if m and cur:
                    info[cur]["spill"] = int(m.group(1)) + int(m.group(2))
                m = re.search(r"Used (\d+) registers", line)
                if m and cur:
                    info[cur]["regs"] = int(m.group(1))
                    s = re.search(r"(\d+) bytes smem", line)
                    info[cur]["smem"] = int(s.group(1)) if s else 0
            sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
            cur = None
            for line in sass.splitlines():
                m = re.search(r"Function : (\S+)", line)
                if m:
                    cur = m.group(1)
                    info.setdefault(cur, {"regs": "?", "spill": 0, "smem": 0})
                    info[cur]["n"], info[cur]["ops"] = 0, {}
                    continue
                m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
                if m and cur:
                    info[cur]["n"] += 1
                    op = m.group(1)
                    for w in WATCH:
                        if op.startswith(w):
                            info[cur]["ops"][w] = info[cur]["ops"].get(w, 0) + 1
            names = demangle(sorted(info))
            for f in sorted(info, key=lambda k: names[k]):
                d = info[f]
                rows.append((src, names[f][:64], d["regs"], d["spill"], d["smem"], d.get("n", 0),
                             ", ".join("%s %d" % kv for kv in sorted(d.get("ops", {}).items(), key=lambda kv: WATCH.index(kv[0])))))
    print("# Static facts per kernel (sm_100a, `nvcc -O3 -lineinfo`, `ptxas -v` + `cuobjdump -sass`; `python tools/static_report.py`)\n")
    print("Counts are static SASS instructions of the whole kernel (all paths), not executed instructions.  tcgen05 = `UTCHMMA` (MMA), "
          "`UTCBAR` (commit), `LDTM` (TMEM load); TMA = `UTMALDG` / `UTMAPF`; `SYNCS` = mbarrier; `LDGSTS` = cp.async; `MUFU` = special-function unit.\n")
    print("| file | kernel | regs | spill B | static smem B | SASS instr | notable mnemonics |")
    print("|---|---|---|---|---|---|---|")
    for r in rows:
        print("| %s | `%s` | %s | %s | %s | %s | %s |" % r)


if __name__ == "__main__":
    main()
