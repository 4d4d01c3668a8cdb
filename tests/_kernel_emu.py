"""TEST INFRASTRUCTURE, not product code: runs the *source text* of the library's CUDA-core kernels on the host.

``build()`` takes whole ``namespace w2l { ... }`` sections out of the ``.cu`` files under wav2letter_pytorch_b200/csrc --
kernels, ``__device__`` helpers, constants and structs exactly as nvcc sees them -- compiles them with g++ against
tests/kernel_emu_runtime.h (a fiber-per-CUDA-thread stand-in for blocks, warps, shared memory, barriers, shuffles and
atomics; CUDA's own vector-type and bf16 headers compile for the host) and adds one launcher per requested kernel, loaded
through ctypes.  Functions written in inline PTX cannot be compiled for the host: the caller names them in ``drop`` and
supplies host replacements in ``extra`` (each replacement cites the PTX it stands for).

What executes is therefore the kernels' own index arithmetic, masks, reductions, synchronisation protocol and roundings;
what is NOT covered is everything device-specific (launch geometry chosen by the C wrappers, alignment faults, memory
ordering, the approximate SFU functions, tensor cores / TMA / TMEM -- conv_gemm.cu is out of reach by construction).

Used by tests/test_kernel_emu*.py.  Nothing under wav2letter_pytorch_b200/ imports this module."""
import ctypes
import hashlib
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "wav2letter_pytorch_b200", "csrc")
HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
# W2L_EMU_ASAN=1: AddressSanitizer in the emulated kernels (run python with LD_PRELOAD=$(gcc -print-file-name=libasan.so) and
# ASAN_OPTIONS=detect_leaks=0): tensors then come from ASan's allocator with red zones around them, so a kernel that reads or
# writes past a buffer -- a vector access over the end of a row, an off-by-one tail -- aborts with a report (memcheck on the host)
SANITIZE = ["-fsanitize=address", "-fno-omit-frame-pointer", "-g"] if os.environ.get("W2L_EMU_ASAN") == "1" else []
# W2L_EMU_UBSAN=1: UndefinedBehaviorSanitizer, above all its alignment check -- CUDA's vector types keep their alignment on the host,
# so a 16-byte load / store through a pointer that is not 16-byte aligned (a misaligned-address fault on the GPU) is reported, as are
# signed overflows and out-of-range shifts in the index arithmetic.  Reports go to stderr and abort (-fno-sanitize-recover).
if os.environ.get("W2L_EMU_UBSAN") == "1":
    SANITIZE += ["-fsanitize=undefined", "-fno-sanitize-recover=all", "-fno-sanitize=vptr,float-cast-overflow,float-divide-by-zero", "-g"]


def _match(text, open_pos, open_ch, close_ch):
    depth = 0
    for i in range(open_pos, len(text)):
        if text[i] == open_ch:
            depth += 1
        elif text[i] == close_ch:
            depth -= 1
            if depth == 0:
                return i
    raise ValueError("unbalanced %s%s" % (open_ch, close_ch))


def _strip_comments(text):
    text = re.sub(r"//[^\n]*", "", text)
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def find_function(text, name):
    """(start, end, parameter text) of the definition of ``name`` in ``text``: from its ``template <...>`` / ``__global__`` /
    ``__device__`` qualifier to the closing brace of its body"""
    for m in re.finditer(r"\b%s\s*\(" % re.escape(name), text):
        start = max(text.rfind(q, 0, m.start()) for q in ("__global__", "__device__", "\nstatic ", "\nint "))
        if start < 0:
            continue
        between = text[start:m.start()]
        if ";" in between or "}" in between or "{" in between:         # a call site, not the definition
            continue
        close_paren = _match(text, m.end() - 1, "(", ")")
        brace = text.index("{", close_paren)
        if text[close_paren + 1:brace].strip():
            continue
        tm = re.search(r"template\s*<[^>]*>\s*$", text[:start])
        if tm:
            start = tm.start()
        return start, _match(text, brace, "{", "}") + 1, text[m.end():close_paren]
    raise KeyError("definition of %s not found" % name)


def namespace_sections(path):
    """the contents of every ``namespace w2l { ... }`` section of a .cu file, concatenated"""
    text = open(path).read()
    out = []
    for m in re.finditer(r"^namespace w2l \{", text, flags=re.M):
        end = _match(text, m.end() - 1, "{", "}")
        out.append(text[m.end():end])
    if not out:
        raise ValueError("no namespace w2l section in " + path)
    return "\n".join(out)


def extern_c_sections(path):
    """the bodies of the ``extern "C" { ... }`` blocks of a .cu file plus its single ``extern "C" T f(...) { ... }`` definitions, as
    one ``extern "C" { ... }`` text: the C-ABI wrappers (argument checks, launch planning) exactly as the library compiles them"""
    text = open(path).read()
    out = []
    for m in re.finditer(r'^extern "C"\s*(\{)?', text, flags=re.M):
        if m.group(1):
            end = _match(text, m.end() - 1, "{", "}")
            out.append(text[m.end():end])
        else:
            brace = text.index("{", m.end())
            end = _match(text, brace, "{", "}")
            out.append(text[m.end():end + 1])
    return 'extern "C" {\n' + "\n".join(out) + "\n}\n"


def _split_top(text):
    parts, depth, cur = [], 0, ""
    for ch in text:
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
            continue
        depth += ch in "(<[{"
        depth -= ch in ")>]}"
        cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def transform_launches(text):
    """``kernel<<<grid, block[, smem[, stream]]>>>(args);`` -> a call of the fiber runtime (the enclosing function returns
    W2L_ERR_CUDA when the emulation reports a deadlock / overrun, as a failed launch would)"""
    out, pos = "", 0
    while True:
        i = text.find("<<<", pos)
        if i < 0:
            return out + text[pos:]
        j = i                                                   # walk back over the kernel name (identifier + template arguments)
        while j > 0 and text[j - 1].isspace():
            j -= 1
        if text[j - 1] == ">":
            depth, j = 0, j - 1
            while True:
                depth += text[j] == ">"
                depth -= text[j] == "<"
                if depth == 0:
                    break
                j -= 1
        while j > 0 and (text[j - 1].isalnum() or text[j - 1] in "_:"):
            j -= 1
        name = text[j:i].strip()
        k = text.index(">>>", i)
        cfg = _split_top(text[i + 3:k])
        op = text.index("(", k)
        cl = _match(text, op, "(", ")")
        semi = text.index(";", cl)
        assert not text[cl + 1:semi].strip(), "unexpected text after a kernel launch"
        smem = cfg[2] if len(cfg) > 2 else "0"
        call = ("{ const dim3 emu_g_(%s), emu_b_(%s); if (emu::launch(emu_g_.x, emu_g_.y, emu_g_.z, emu_b_.x, emu_b_.y, emu_b_.z, (size_t)(%s), "
                "[=]() { %s(%s); })) { w2l::set_error(\"emulated launch of %s failed\"); return W2L_ERR_CUDA; } }"
                % (cfg[0], cfg[1], smem, name, text[op + 1:cl], name.replace('"', "")))
        out += text[pos:j] + call
        pos = semi + 1


_CTYPES = (("int64_t", ctypes.c_int64), ("uint64_t", ctypes.c_uint64), ("int32_t", ctypes.c_int32), ("uint32_t", ctypes.c_uint32),
           ("size_t", ctypes.c_size_t), ("double", ctypes.c_double), ("float", ctypes.c_float), ("unsigned", ctypes.c_uint),
           ("int", ctypes.c_int))


def _params(param_text):
    out, depth, cur = [], 0, ""
    for ch in param_text + ",":
        if ch == "," and depth == 0:
            p = " ".join(cur.split())
            cur = ""
            if not p:
                continue
            name = re.findall(r"[A-Za-z_][A-Za-z_0-9]*", p)[-1]
            call = name
            if "*" in p:
                ct = ctypes.c_void_p
            else:
                hit = [c for key, c in _CTYPES if re.search(r"\b%s\b" % key, p)]
                if hit:
                    ct = hit[0]
                else:                          # a struct passed by value: the launcher takes its address (ctypes.addressof / data_ptr)
                    ct = ctypes.c_void_p
                    p = "const %s* %s" % (p[:p.rindex(name)].strip(), name)
                    call = "*" + name
            out.append((p, call, ct))
            continue
        depth += ch in "(<["
        depth -= ch in ")>]"
        cur += ch
    return out


class EmuError(RuntimeError):
    pass


class Emu:
    def __init__(self, lib, sigs):
        self._lib, self._sigs = lib, sigs

    def launch(self, kernel, grid, block, *args, smem=0):
        """as ``kernel<<<grid, block, smem>>>(args...)``: grid / block are ints or up-to-3 tuples, pointers are ints (data_ptr())"""
        g = tuple(grid) if isinstance(grid, (tuple, list)) else (grid,)
        b = tuple(block) if isinstance(block, (tuple, list)) else (block,)
        g, b = g + (1,) * (3 - len(g)), b + (1,) * (3 - len(b))
        sym, sig = self._sigs[kernel]
        assert len(args) == len(sig), "%s takes %d arguments, got %d" % (kernel, len(sig), len(args))
        conv = [ctypes.c_void_p(a or 0) if ct is ctypes.c_void_p else ct(a) for a, (_, _, ct) in zip(args, sig)]
        fn = getattr(self._lib, sym)
        fn.restype = ctypes.c_int
        rc = fn(*(ctypes.c_int(int(v)) for v in g + b), ctypes.c_size_t(int(smem)), *conv)
        if rc:
            raise EmuError("%s: %s" % (kernel, {1: "deadlock: every live thread waits at a barrier / warp collective that cannot complete",
                                                2: "a polling loop did not finish within the yield budget",
                                                3: "a block wrote past the end of its dynamic shared memory"}.get(rc, "error %d" % rc)))


def build(sources, kernels, drop=(), extra="", post="", helpers_from_common=("pack_bf16x2",), subs=(), c_abi=False, opt="-O1"):
    """sources: .cu file names under csrc/; kernels: kernel names (``name<float>`` instantiates a template); drop: functions cut
    from the sources (inline PTX, host functions that launch kernels); extra: C++ placed inside namespace w2l ahead of the sources
    (host replacements for dropped functions); post: C++ appended at file scope (e.g. extern "C" access to host planning code);
    drop may be a dict name -> replacement text (put where the function stood); subs: (regex, replacement) pairs applied to the
    source text (inline PTX statements inside a kernel body); c_abi: also compile the files' extern "C" wrappers, with their
    ``<<<...>>>`` launches turned into calls of the fiber runtime -- the result exports the library's own entry points"""
    common = open(os.path.join(CSRC, "common.cuh")).read()
    parts = ['#include "kernel_emu_runtime.h"', "namespace w2l {"]
    for h in helpers_from_common:
        s, e, _ = find_function(common, h)
        parts.append(common[s:e])
    parts.append(extra)
    body = "\n".join(namespace_sections(os.path.join(CSRC, f)) for f in sources)
    for name in drop:
        s, e, _ = find_function(body, name)
        body = body[:s] + (drop[name] if isinstance(drop, dict) else "") + body[e:]
    for pat, rep in subs:
        body, n_sub = re.subn(pat, rep, body, flags=re.S)
        if n_sub == 0:
            raise ValueError("substitution %r matched nothing" % pat)
    body = transform_launches(body)
    # dynamic shared memory: `extern __shared__ [__align__(n)] T name[];` -> a typed view of the launch's buffer
    body = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][A-Za-z_0-9:]*)\s+([A-Za-z_][A-Za-z_0-9]*)\s*\[\s*\]\s*;",
                  r"\1* \2 = reinterpret_cast<\1*>(emu::dyn_smem);", body)
    left = re.findall(r"\basm\b", _strip_comments(body))
    if left:
        raise ValueError("inline PTX left in %s: name the functions that hold it in `drop` and replace them in `extra`" % (sources,))
    parts.append(body)
    parts.append("}  // namespace w2l")
    parts.append("using namespace w2l;       // launcher signatures name the library's own types")
    if c_abi:
        parts.extend(transform_launches(extern_c_sections(os.path.join(CSRC, f))) for f in sources)
    parts.append(post)
    sigs = {}
    for spec in kernels:
        m = re.match(r"([A-Za-z_0-9]+)(?:<(.*)>)?$", spec)
        name, targs = m.group(1), m.group(2)
        s, e, ptext = find_function(body, name)
        if "__global__" not in body[s:e].split("(")[0]:
            raise ValueError("%s is not a __global__ function" % name)
        if targs:
            tnames = re.findall(r"(?:typename|class|int|bool)\s+([A-Za-z_][A-Za-z_0-9]*)", re.match(r"template\s*<([^>]*)>", body[s:e]).group(1))
            for tn, ta in zip(tnames, [t.strip() for t in targs.split(",")]):
                ptext = re.sub(r"\b%s\b" % tn, ta, ptext)
        ps = _params(ptext)
        sym = "emu_" + re.sub(r"[^A-Za-z0-9_]", "_", spec)
        sigs[spec] = (sym, ps)
        parts.append("""
extern "C" int %s(int emu_gx, int emu_gy, int emu_gz, int emu_bx, int emu_by, int emu_bz, size_t emu_smem%s) {
  return emu::launch(emu_gx, emu_gy, emu_gz, emu_bx, emu_by, emu_bz, emu_smem, [=]() { w2l::%s(%s); });
}
""" % (sym, "".join(", " + p for p, _, _ in ps), spec, ", ".join(n for _, n, _ in ps)))
    code = "\n".join(parts)
    runtime = open(os.path.join(HERE, "kernel_emu_runtime.h")).read()
    tag = hashlib.sha1((code + runtime + " ".join(SANITIZE)).encode()).hexdigest()[:16]
    out_dir = os.path.join(tempfile.gettempdir(), "w2l_kernel_emu")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "emu_%s.so" % tag)
    if not os.path.exists(so):
        cpp = os.path.join(out_dir, "emu_%s.cpp" % tag)
        with open(cpp, "w") as fh:
            fh.write(code)
        tmp = so + ".%d.tmp" % os.getpid()
        r = subprocess.run(["g++", opt, "-w", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I" + CUDA_INC, "-I" + HERE] + SANITIZE
                           + [cpp, "-o", tmp], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("host compilation of the kernel sources failed (%s):\n%s" % (cpp, r.stderr[-6000:]))
        os.replace(tmp, so)
    emu = Emu(ctypes.CDLL(so), sigs)
    emu.lib = emu._lib
    return emu


def available():
    return os.path.exists(os.path.join(CUDA_INC, "cuda_bf16.h")) and subprocess.run(["which", "g++"], capture_output=True).returncode == 0
