"""ctypes binding of libw2l_sm100.so (C ABI declared in include/w2l_sm100.h) + the in-tree nvcc build.

There is deliberately NO CPU fallback here: if the shared library is missing or a call fails the caller
gets a RuntimeError carrying ``w2l_last_error()``.
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libw2l_sm100.so")
SOURCES = ["runtime.cu", "decode.cu", "ctc.cu", "conv_gemm.cu", "elementwise.cu", "novograd.cu", "metrics.cu", "depthwise.cu", "comm.cu", "features.cu", "beam_search.cu"]
_lock = threading.Lock()
_lib = None


_HASH_TAG = b"W2L_SRC_HASH="


def source_hash():
    """sha256 (16 hex digits) over csrc/* and include/w2l_sm100.h: what the shared library was built from"""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    files.append(os.path.join(os.path.dirname(_HERE), "include", "w2l_sm100.h"))
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def built_hash(path=None):
    """the source hash embedded in an existing shared library (read from the file, without loading it), or None"""
    path = path or LIB_PATH
    try:
        with open(path, "rb") as fh:
            blob = fh.read()
    except OSError:
        return None
    i = blob.find(_HASH_TAG)
    return blob[i + len(_HASH_TAG): i + len(_HASH_TAG) + 16].decode(errors="replace") if i >= 0 else None


def _needs_build():
    """a library is current only if it was built from exactly the sources in the tree (the prebuilt .so travels to the GPU box,
    and the number measured there must come from HEAD's kernels)"""
    return built_hash() != source_hash()


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into an in-tree shared library (nvcc cross-compiles without a GPU)."""
    if not force and not _needs_build():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    builddir = os.path.join(_HERE, "build")
    os.makedirs(builddir, exist_ok=True)
    procs = []
    src_hash = source_hash()
    for src in SOURCES:
        obj = os.path.join(builddir, src.replace(".cu", ".o"))
        objs.append(obj)
        srcp = os.path.join(CSRC, src)
        if not force and src != "runtime.cu" and os.path.exists(obj) and os.path.getmtime(obj) > max(
                os.path.getmtime(srcp), os.path.getmtime(os.path.join(CSRC, "common.cuh")),
                os.path.getmtime(os.path.join(os.path.dirname(_HERE), "include", "w2l_sm100.h"))):
            continue
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
               "-DW2L_SRC_HASH=\"%s\"" % src_hash, "-c", srcp, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode()))
        if verbose and out:
            print(out.decode())
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout.decode()))
    os.replace(tmp, LIB_PATH)                   # a new inode: a process that has the old library mapped keeps a valid mapping
    if built_hash() != src_hash:
        raise RuntimeError("libw2l_sm100.so does not carry the hash of the sources it was just built from")
    return LIB_PATH


c_i32, c_i64, c_f32, c_u64 = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64
c_ptr, c_size = ctypes.c_void_p, ctypes.c_size_t


class ConvDesc(ctypes.Structure):
    """Mirror of w2l_conv_desc."""
    _fields_ = [(n, c_i32) for n in ("B", "T_out", "Cin", "Cout", "Cout_pad", "k", "dilation", "x_rows", "x_row_offset",
                                     "y_rows", "y_row_offset", "ldy", "y_dtype", "act", "x_dtype")]


class BnReduce(ctypes.Structure):
    """Mirror of w2l_bn_reduce: the BatchNorm-backward reduction folded into the epilogue of w2l_conv1d_dgrad_wt_bnred."""
    _fields_ = [("z", ctypes.c_void_p), ("drop_mask", ctypes.c_void_p), ("scale", ctypes.c_void_p), ("shift", ctypes.c_void_p),
                ("mean", ctypes.c_void_p), ("lens", ctypes.c_void_p), ("red", ctypes.c_void_p), ("B", c_i32), ("T", c_i32),
                ("pad_left", c_i32), ("pad_right", c_i32), ("act", c_i32), ("drop_p", ctypes.c_float)]


# name -> (restype, argtypes); every symbol declared in include/w2l_sm100.h appears here
SIGNATURES = {
    "w2l_version": (c_i32, []),
    "w2l_last_error": (ctypes.c_char_p, []),
    "w2l_launch_count": (c_i64, []),
    "w2l_source_hash": (ctypes.c_char_p, []),
    "w2l_set_sm_budget": (c_i32, [c_i32]),
    "w2l_get_sm_budget": (c_i32, []),
    "w2l_edit_distance_host": (c_i64, [c_ptr, c_i64, c_ptr, c_i64]),
    "w2l_edit_distance_batch_host": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_i32]),
    "w2l_greedy_decode_workspace_bytes": (c_size, [c_i64, c_i64]),
    "w2l_greedy_decode": (c_i32, [c_ptr, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr, c_i32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                  c_size, c_ptr]),
    "w2l_string_metrics_workspace_bytes": (c_size, [c_i64, c_i64, c_i64]),
    "w2l_string_metrics": (c_i32, [c_ptr, c_ptr, c_i64, c_i64, c_i32, c_ptr, c_ptr, c_i64, c_f32, c_f32, c_f32, c_ptr, c_ptr, c_size,
                                   c_ptr]),
    "w2l_ctc_loss_workspace_bytes": (c_size, [c_i64, c_i64, c_i64]),
    "w2l_ctc_loss": (c_i32, [c_ptr, c_i32, c_i64, c_i64, c_i64, c_i64, c_i64, c_ptr, c_i64, c_ptr, c_ptr, c_i32, c_i32, c_i32,
                             c_ptr, c_ptr, c_ptr, c_ptr, c_size, c_ptr]),
    "w2l_conv1d_fwd": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, ctypes.POINTER(ConvDesc), c_ptr]),
    "w2l_set_gemm_scratch": (c_i32, [c_ptr, c_size]),
    "w2l_set_dropout_epoch": (c_i32, [c_ptr]),
    "w2l_conv1d_fwd_tail_parts": (c_i32, [ctypes.POINTER(ConvDesc)]),
    "w2l_conv1d_dgrad": (c_i32, [c_ptr, c_ptr, c_ptr, ctypes.POINTER(ConvDesc), c_ptr]),
    "w2l_conv1d_dgrad_wt": (c_i32, [c_ptr, c_ptr, c_ptr, ctypes.POINTER(ConvDesc), c_ptr]),
    "w2l_pack_wt": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_conv1d_wgrad_splits": (c_i32, [ctypes.POINTER(ConvDesc)]),
    "w2l_conv1d_wgrad": (c_i32, [c_ptr, c_ptr, c_ptr, ctypes.POINTER(ConvDesc), c_ptr]),
    "w2l_im2col_ncw_f32": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr]),
    "w2l_reflect_halo_f32": (c_i32, [c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_log_softmax_bwd_f32": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i64, c_i32, c_i32, c_ptr]),
    "w2l_colsum_f32": (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "w2l_pack_wt_f32": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_depthwise_fwd_f32": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 8 + [c_ptr, c_ptr]),
    "w2l_depthwise_dgrad_f32": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 7 + [c_ptr, c_ptr]),
    "w2l_depthwise_dgrad_strided_f32": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 8 + [c_ptr, c_ptr]),
    "w2l_depthwise_wgrad_f32": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 8 + [c_ptr, c_ptr]),
    "w2l_conv1d_dgrad_wt_bnred": (c_i32, [c_ptr, c_ptr, c_ptr, ctypes.POINTER(ConvDesc), ctypes.POINTER(BnReduce), c_ptr]),
    "w2l_conv1d_wgrad_t": (c_i32, [c_ptr, c_i64, c_ptr, c_i64, c_ptr, ctypes.POINTER(ConvDesc), c_ptr]),
    "w2l_tm_to_ct_f32": (c_i32, [c_ptr, c_ptr] + [c_i32] * 7 + [c_ptr]),
    "w2l_depthwise_fwd": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 8 + [c_ptr, c_ptr]),
    "w2l_depthwise_dgrad": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 7 + [c_ptr, c_ptr]),
    "w2l_depthwise_dgrad_strided": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 8 + [c_ptr, c_ptr]),
    "w2l_depthwise_wgrad": (c_i32, [c_ptr, c_ptr, c_ptr] + [c_i32] * 8 + [c_ptr, c_ptr]),
    "w2l_im2col_ncw": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr]),
    "w2l_im2col_tm": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_col2im_tm": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_tm_to_ncw": (c_i32, [c_ptr, c_i32, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_ncw_to_tm": (c_i32, [c_ptr, c_ptr, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_bn_stats": (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_ptr]),
    "w2l_bn_finalize": (c_i32, [c_ptr, c_i64, c_i32, c_ptr, c_ptr, c_ptr, c_f32, c_f32, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr,
                                c_ptr, c_ptr]),
    "w2l_lens_chain": (c_i32, [c_ptr, c_i32, c_i32, c_ptr, c_i32, c_ptr, c_ptr, c_ptr]),
    "w2l_bn_act_pad": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_f32,
                               c_u64, c_ptr, c_ptr, c_ptr]),
    "w2l_reflect_halo": (c_i32, [c_ptr, c_i32, c_i32, c_i32, c_i32, c_i32, c_ptr]),
    "w2l_bn_act_bwd_reduce": (c_i32, [c_ptr] * 10 + [c_i32] * 6 + [c_f32, c_u64, c_ptr, c_ptr, c_ptr]),
    "w2l_bn_finalize_act_pad": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_ptr, c_f32, c_f32] + [c_ptr] * 8 + [c_i32] * 6
                                + [c_f32, c_u64, c_ptr, c_ptr, c_ptr, c_i32, c_ptr]),
    "w2l_bn_act_bwd_apply": (c_i32, [c_ptr] * 12 + [c_i32, c_ptr] + [c_i32] * 6 + [c_f32, c_u64, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_ptr]),
    "w2l_log_softmax": (c_i32, [c_ptr, c_i32, c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "w2l_log_softmax_bwd": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i64, c_i32, c_i32, c_ptr]),
    "w2l_colsum": (c_i32, [c_ptr, c_i64, c_i32, c_i32, c_ptr, c_ptr]),
    "w2l_cast_bf16": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr]),
    "w2l_prefix_beam_search_host": (c_i32, [c_ptr, c_ptr, c_i64, c_i64, c_i64, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_i32, ctypes.c_double,
                                            ctypes.c_double, ctypes.c_double, c_ptr, c_ptr, c_ptr, c_i64, c_ptr, c_ptr, c_i32]),
    "w2l_logmel_workspace_bytes": (c_size, [c_i32, c_i32, c_i32]),
    "w2l_logmel_features": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr, c_ptr, c_i32, c_f32, c_f32, c_f32, c_f32,
                                    c_ptr, c_i32, c_ptr, c_size, c_ptr]),
    "w2l_grad_allreduce": (c_i32, [c_ptr, c_ptr, c_ptr, c_i64, c_i64, c_i32, c_i32, ctypes.c_uint32, c_i32, c_ptr]),
    "w2l_novograd_chunk": (c_i32, []),
    "w2l_novograd_step": (c_i32, [c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_f32, c_f32, c_f32, c_f32, c_f32,
                                  c_i32, c_ptr, c_ptr]),
}


def load():
    """Returns the loaded library (building it first if the .so is missing)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                build()
            elif os.environ.get("W2L_ALLOW_STALE_LIB", "0") != "1" and _needs_build():
                # never run (or measure) kernels that are not the ones in the tree: rebuild where nvcc exists, else fail loudly
                try:
                    build()
                except Exception as e:  # noqa: BLE001
                    raise RuntimeError("libw2l_sm100.so was built from other sources (%s) than the tree holds (%s) and could not be "
                                       "rebuilt here: %s" % (built_hash(), source_hash(), e)) from e
            lib = ctypes.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


class W2LError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().w2l_last_error().decode(errors="replace")
        raise W2LError("libw2l_sm100 %s failed (code %d): %s" % (what, rc, msg))


def launch_count():
    return int(load().w2l_launch_count())
