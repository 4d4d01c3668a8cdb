"""TEST INFRASTRUCTURE ONLY -- libw2l_sm100's C ABI answered on the host by the library's own sources.

tests/_emu_backend.py repeats each C wrapper's launch arithmetic in Python (and poisons every output buffer).  This module
goes one level lower: every ``.cu`` file is compiled for the host TOGETHER WITH its ``extern "C"`` wrappers
(tests/_kernel_emu.py, ``c_abi=True``: ``<<<...>>>`` launches become calls of the fiber runtime), and ``EmuLibrary`` hands out
those entry points under the names and ctypes signatures of ``wav2letter_pytorch_b200._lib.SIGNATURES``.  With
``install(monkeypatch)`` the UNMODIFIED ``functional.py`` -- argument marshalling, workspace sizing, descriptor structs -- and
everything above it run on CPU tensors: what a GPU run exercises, minus the GPU.

Host-only entry points (edit distance, prefix beam search, version) come from the real shared library, which loads without
a GPU; ``w2l_grad_allreduce`` (NVLink multimem) is not emulated.  Nothing under wav2letter_pytorch_b200/ imports this module."""
import contextlib
import ctypes
import functools

import torch

import _emu_backend as E
import _kernel_emu as KE

_POST = r"""
extern "C" const char* emu_last_error() { return w2l::g_err; }
extern "C" long long emu_launch_count() { return w2l::g_launches; }
extern "C" void emu_set_sm_budget(int sms) { w2l::g_sm_budget = sms; }
"""
_CTC_DROP = ["fast_ex2", "fast_lg2", "cp_async4", "cp_async16", "cp_async_commit", "cp_async_wait", "slot_put", "slot_load"]


@functools.lru_cache(maxsize=None)
def builds():
    out = [KE.build([f], [], c_abi=True, post=_POST) for f in ("decode.cu", "depthwise.cu", "novograd.cu", "metrics.cu", "features.cu")]
    out.append(KE.build(["elementwise.cu"], [], c_abi=True, post=_POST, drop=E.ELEMENTWISE_DROP, extra=E.ELEMENTWISE_PTX))
    out.append(KE.build(["ctc.cu"], [], c_abi=True, post=_POST, drop=_CTC_DROP, extra=E.CTC_PTX))
    out.append(KE.build(["conv_gemm.cu"], [], helpers_from_common=("pack_bf16x2", "make_smem_desc", "make_idesc_bf16"), subs=E.GEMM_SUBS,
                        drop=E.GEMM_DROP, c_abi=True, post=_POST, opt="-O2"))
    return out


class EmuLibrary:
    """looks like the ctypes.CDLL that ``_lib.load()`` returns"""

    def __init__(self):
        from wav2letter_pytorch_b200 import _lib
        self._parts = [b.lib for b in builds()]
        for p in self._parts:
            p.emu_last_error.restype = ctypes.c_char_p
            p.emu_launch_count.restype = ctypes.c_longlong
        self._real = ctypes.CDLL(_lib.LIB_PATH) if _lib.os.path.exists(_lib.LIB_PATH) else None
        self._err = b""
        self.emulated, self.host_only, self.missing = [], [], []
        for name, (res, args) in _lib.SIGNATURES.items():
            if name in ("w2l_last_error", "w2l_launch_count", "w2l_set_sm_budget", "w2l_get_sm_budget"):
                continue
            owner = next((p for p in self._parts if hasattr(p, name)), None)
            if owner is not None:
                fn = getattr(owner, name)
                fn.restype, fn.argtypes = res, args
                setattr(self, name, self._wrap(fn, owner) if res is ctypes.c_int32 else fn)
                self.emulated.append(name)
            elif self._real is not None and hasattr(self._real, name):
                fn = getattr(self._real, name)
                fn.restype, fn.argtypes = res, args
                setattr(self, name, fn)
                self.host_only.append(name)
            else:
                self.missing.append(name)

    def _wrap(self, fn, owner):
        def call(*a):
            rc = fn(*a)
            if rc:
                self._err = owner.emu_last_error()
            return rc
        return call

    def w2l_last_error(self):
        return self._err

    def w2l_launch_count(self):
        return sum(int(p.emu_launch_count()) for p in self._parts)

    def w2l_set_sm_budget(self, sms):
        for p in self._parts:
            p.emu_set_sm_budget(int(sms))
        return 0

    def w2l_get_sm_budget(self):
        return 148


@functools.lru_cache(maxsize=None)
def library():
    return EmuLibrary()


class _TorchProxy:
    """``torch`` as functional.py sees it, except that byte buffers (workspaces, the GEMM scratch) start on a 256-byte boundary,
    as cudaMalloc'd memory does and as the C wrappers require (the CPU allocator only guarantees 64)"""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def _aligned(make, shape, kw):
        n = int(shape[0]) if isinstance(shape, (tuple, list, torch.Size)) else int(shape)
        raw = make((n + 256,), **kw)
        off = (-raw.data_ptr()) % 256
        return raw[off:off + n]

    def empty(self, *shape, **kw):
        """``torch.empty`` as functional.py calls it for outputs and workspaces -- POISONED (NaN / 0xFF bytes / a large negative
        integer), so that a kernel which reads memory nobody wrote, or a wrapper that leaves part of an output unwritten, shows up in
        the results instead of depending on what the allocator happened to hand back (compute-sanitizer's initcheck, on the host)"""
        if kw.get("dtype") is torch.uint8 and len(shape) == 1:
            t = self._aligned(torch.empty, shape[0], kw)
        else:
            t = torch.empty(*shape, **kw)
        if t.is_floating_point():
            t.fill_(float("nan"))
        elif t.dtype is torch.uint8:
            t.fill_(0xFF)
        elif t.dtype in (torch.int32, torch.int64, torch.int16):
            t.fill_(-0x01010102)
        return t

    def empty_like(self, t, **kw):
        out = torch.empty_like(t, **kw)
        if out.is_floating_point():
            out.fill_(float("nan"))
        elif out.dtype is not torch.bool:
            out.fill_(-0x01010102 if out.dtype is not torch.uint8 else 0xFF)
        return out

    def zeros(self, *shape, **kw):
        if kw.get("dtype") is torch.uint8 and len(shape) == 1:
            return self._aligned(torch.zeros, shape[0], kw)
        return torch.zeros(*shape, **kw)


_keep = []


def install(monkeypatch):
    """``_lib.load()`` -> the emulated library; functional.py's three CUDA-isms (device check, current stream, device guard) neutralised"""
    from wav2letter_pytorch_b200 import _lib, layers
    from wav2letter_pytorch_b200 import functional as F
    lib = library()
    monkeypatch.setattr(_lib, "load", lambda: lib)
    monkeypatch.setattr(F, "_need_cuda", lambda *ts: None)
    monkeypatch.setattr(F, "_stream", lambda: None)
    monkeypatch.setattr(F, "to_cuda", lambda t, who=None: t)
    monkeypatch.setattr(F, "torch", _TorchProxy())
    _keep.append({})                                   # the scratch registered with the emulated library must outlive the test
    monkeypatch.setattr(F, "_gemm_scratch", _keep[-1])
    monkeypatch.setattr(torch.cuda, "device", lambda *a, **k: contextlib.nullcontext())
    monkeypatch.setattr(layers.WgradStream, "enabled", False)
    monkeypatch.setattr(layers.FusedBnReduce, "enabled", True)       # the opt-in fused BatchNorm-backward reduction: covered here on the host
    return F
