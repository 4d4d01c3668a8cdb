#!/usr/bin/env python
"""Which kernels did a source change touch?  Compiles csrc/<file>.cu of two git revisions for sm_100a (nvcc cross-compiles without a
GPU) and compares the SASS of every kernel both contain -- the check behind "this change does not touch the measured default path"
when no GPU is at hand.  Mnemonic-level comparison (`--loose`) ignores register numbers and immediates.

    python tools/sass_diff.py bf3ade6 HEAD ctc.cu depthwise.cu
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = "wav2letter_pytorch_b200/csrc"


def build(rev, name, tmp):
    d = os.path.join(tmp, rev.replace("/", "_"))                  # the revision's files in the repo's own layout (common.cuh includes ../../include)
    src_dir, inc_dir = os.path.join(d, CSRC), os.path.join(d, "include")
    os.makedirs(src_dir, exist_ok=True)
    os.makedirs(inc_dir, exist_ok=True)
    for path, dst in ((CSRC + "/" + name, src_dir), (CSRC + "/common.cuh", src_dir), ("include/w2l_sm100.h", inc_dir)):
        text = subprocess.run(["git", "show", "%s:%s" % (rev, path)], cwd=ROOT, capture_output=True, text=True, check=True).stdout
        with open(os.path.join(dst, os.path.basename(path)), "w") as fh:
            fh.write(text)
    obj = os.path.join(src_dir, name.replace(".cu", ".o"))
    subprocess.run([os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
                    "-std=c++17", "-Xcompiler", "-fPIC", "-c", os.path.join(src_dir, name), "-o", obj], check=True)
    return obj


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1).strip())
    return funcs


def demangle(names):
    if not names:
        return {}
    out = subprocess.run(["c++filt"] + list(names), capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout.splitlines()
    return dict(zip(names, (re.sub(r"\(.*", "", o)[:70] for o in out)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rev_a")
    ap.add_argument("rev_b")
    ap.add_argument("files", nargs="+")
    ap.add_argument("--loose", action="store_true", help="compare mnemonics only (ignore registers and immediates)")
    args = ap.parse_args()
    norm = (lambda L: [re.sub(r"\bU?R\d+\b|\bU?P\d\b", "r", re.sub(r"0x[0-9a-f]+", "imm", x)) for x in L]) if args.loose else (lambda L: L)
    rc = 0
    with tempfile.TemporaryDirectory() as tmp:
        for name in args.files:
            a, b = sass(build(args.rev_a, name, tmp)), sass(build(args.rev_b, name, tmp))
            pretty = demangle(sorted(set(a) | set(b)))
            for f in sorted(set(a) | set(b)):
                if f not in a:
                    state = "new"
                elif f not in b:
                    state = "gone"
                elif norm(a[f]) == norm(b[f]):
                    state = "same"
                else:
                    state = "DIFFERENT (%d -> %d instructions)" % (len(a[f]), len(b[f]))
                    rc = 1
                print("%-14s %-72s %s" % (name, pretty[f], state))
    return rc


if __name__ == "__main__":
    sys.exit(main())
