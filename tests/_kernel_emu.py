"""TEST INFRASTRUCTURE, not product code: runs the *source text* of thread-independent CUDA kernels on the host.

A kernel whose threads never talk to each other (no shared memory, barriers, shuffles or atomics) is a plain C++ function of
(blockIdx, threadIdx).  ``build()`` cuts such kernels -- and the ``__device__`` helpers they call -- verbatim out of the
``.cu`` / ``.cuh`` files under wav2letter_pytorch_b200/csrc, puts a small shim in front (thread-local blockIdx / threadIdx,
``__ldg``; the vector types and bf16 conversions are CUDA's own headers, which compile for the host), adds one launcher per
kernel that walks the grid sequentially, compiles the lot with g++ and loads it through ctypes.  The index arithmetic, masks,
tap loops and bf16 roundings that execute are therefore exactly the ones nvcc compiles for sm_100a; what is NOT covered is
everything device-specific (launch geometry computed by the C wrappers, alignment faults, memory-model effects).

Used by tests/test_kernel_emu.py so that kernels added when no GPU session was available are still executed, against the
oracle, before they first meet hardware.  Nothing under wav2letter_pytorch_b200/ imports this module."""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "wav2letter_pytorch_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"

SHIM = r"""
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cmath>
#include <algorithm>
using std::min;
using std::max;
#define __launch_bounds__(...)
static thread_local uint3 emu_tid, emu_bid;
static thread_local dim3 emu_bdim, emu_gdim;
#define threadIdx emu_tid
#define blockIdx emu_bid
#define blockDim emu_bdim
#define gridDim emu_gdim
template <typename T> static inline T __ldg(const T* p) { return *p; }
#define W2L_PAD_ZERO 0
#define W2L_PAD_REFLECT 1
"""

FORBIDDEN = ("__shared__", "__syncthreads", "__shfl", "atomic", "__syncwarp", "__ballot", "asm volatile", "asm(")


def _match_brace(text, open_pos):
    depth = 0
    for i in range(open_pos, len(text)):
        if text[i] == "{":
            depth += 1
        elif text[i] == "}":
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced braces")


def extract_function(path, name):
    """the full definition (qualifiers, signature, body) of the function called ``name`` in ``path``"""
    text = open(path).read()
    for m in re.finditer(r"\b%s\s*\(" % re.escape(name), text):
        # walk back to the start of the declaration: the nearest preceding __global__ / __device__ qualifier
        starts = [text.rfind(q, 0, m.start()) for q in ("__global__", "__device__")]
        start = max(starts)
        if start < 0:
            continue
        between = text[start:m.start()]
        if ";" in between or "}" in between or "{" in between:     # a call site, not the definition
            continue
        depth, i = 1, m.end()
        while depth:                                              # parameter lists may hold parentheses, e.g. float (&v)[8]
            if text[i] == "(":
                depth += 1
            elif text[i] == ")":
                depth -= 1
            i += 1
        close_paren = i - 1
        brace = text.index("{", close_paren)
        if text[close_paren + 1:brace].strip():
            continue
        end = _match_brace(text, brace)
        return text[start:end + 1], text[m.end():close_paren]
    raise KeyError("%s not found in %s" % (name, path))


_CTYPES = (("int64_t", ctypes.c_int64), ("uint64_t", ctypes.c_uint64), ("int32_t", ctypes.c_int32), ("uint32_t", ctypes.c_uint32),
           ("float", ctypes.c_float), ("int", ctypes.c_int))


def _params(param_text):
    out = []
    for p in param_text.split(","):
        p = " ".join(p.split())
        name = re.findall(r"[A-Za-z_][A-Za-z_0-9]*", p)[-1]
        if "*" in p:
            ct = ctypes.c_void_p
        else:
            ct = next(c for key, c in _CTYPES if re.search(r"\b%s\b" % key, p))
        out.append((p, name, ct))
    return out


class Emu:
    def __init__(self, lib, sigs):
        self._lib, self._sigs = lib, sigs

    def launch(self, kernel, grid, block, *args):
        """grid / block: int or up-to-3 tuples, as in ``kernel<<<grid, block>>>(args...)``; pointers are ints (data_ptr())"""
        g = tuple(grid) if isinstance(grid, (tuple, list)) else (grid,)
        b = tuple(block) if isinstance(block, (tuple, list)) else (block,)
        g, b = g + (1,) * (3 - len(g)), b + (1,) * (3 - len(b))
        sig = self._sigs[kernel]
        assert len(args) == len(sig), "%s takes %d arguments" % (kernel, len(sig))
        conv = [ct(a if a is not None else 0) if ct is not ctypes.c_void_p else ctypes.c_void_p(a or 0) for a, (_, _, ct) in zip(args, sig)]
        getattr(self._lib, "emu_" + kernel)(*(ctypes.c_int(v) for v in g + b), *conv)


def build(kernels, helpers=()):
    """kernels: [(file, kernel_name)], helpers: [(file, device_function_name)] in dependency order"""
    parts, sigs = [SHIM], {}
    for f, name in helpers:
        src, _ = extract_function(os.path.join(CSRC, f), name)
        parts.append(src)
    for f, name in kernels:
        src, ptext = extract_function(os.path.join(CSRC, f), name)
        bad = [w for w in FORBIDDEN if w in src]
        if bad:
            raise ValueError("%s is not thread-independent (%s): it cannot be emulated by a sequential walk" % (name, bad))
        ps = _params(ptext)
        sigs[name] = ps
        parts.append(src)
        parts.append("""
extern "C" void emu_%s(int emu_gx, int emu_gy, int emu_gz, int emu_bx, int emu_by, int emu_bz, %s) {
  emu_gdim = dim3(emu_gx, emu_gy, emu_gz);
  emu_bdim = dim3(emu_bx, emu_by, emu_bz);
  for (unsigned emu_z = 0; emu_z < (unsigned)emu_gz; ++emu_z) for (unsigned emu_y = 0; emu_y < (unsigned)emu_gy; ++emu_y)
    for (unsigned emu_x = 0; emu_x < (unsigned)emu_gx; ++emu_x)
      for (unsigned emu_c = 0; emu_c < (unsigned)emu_bz; ++emu_c) for (unsigned emu_b = 0; emu_b < (unsigned)emu_by; ++emu_b)
        for (unsigned emu_a = 0; emu_a < (unsigned)emu_bx; ++emu_a) {
      emu_bid = make_uint3(emu_x, emu_y, emu_z);
      emu_tid = make_uint3(emu_a, emu_b, emu_c);
      %s(%s);
    }
}
""" % (name, ", ".join(p for p, _, _ in ps), name, ", ".join(n for _, n, _ in ps)))
    code = "\n".join(parts)
    tag = hashlib.sha1(code.encode()).hexdigest()[:16]
    out_dir = os.path.join(tempfile.gettempdir(), "w2l_kernel_emu")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "emu_%s.so" % tag)
    if not os.path.exists(so):
        cpp = os.path.join(out_dir, "emu_%s.cpp" % tag)
        with open(cpp, "w") as fh:
            fh.write(code)
        tmp = so + ".%d.tmp" % os.getpid()
        r = subprocess.run(["g++", "-O1", "-w", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I" + CUDA_INC, cpp, "-o", tmp],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host compilation of the kernel sources failed:\n" + r.stderr[-4000:])
        os.replace(tmp, so)
    return Emu(ctypes.CDLL(so), sigs)


def available():
    return os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")) and subprocess.run(["which", "g++"], capture_output=True).returncode == 0
