"""TEST INFRASTRUCTURE ONLY -- run the `-m gpu` test files on a machine WITHOUT a GPU:

    python -m pytest tests -m gpu --emulate-gpu

``enable()`` (called from tests/conftest.py when the flag is given) makes "cuda" mean "these CPU tensors" for one pytest session:

* a ``TorchFunctionMode`` maps ``device="cuda"`` / ``.cuda()`` / ``.to("cuda")`` to the CPU, answers ``Tensor.is_cuda`` with True and
  turns ``pin_memory`` / ``record_stream`` into no-ops;
* ``torch.cuda`` queries the tests and the host package make (availability, current device / stream, streams, events, device
  guards, synchronize, memory info) get inert stand-ins;
* ``wav2letter_pytorch_b200._lib.load()`` returns the emulated library of tests/_emu_cabi.py -- the C wrappers and kernels compiled
  for the host from the library's own sources.

So the GPU tests' own code (fixtures, shapes, tolerances), the UNMODIFIED host package including its stream choreography (with inert
streams), and every kernel's source execute; only the hardware is missing.  Purpose: a GPU test written when no GPU session was
left must not meet the hardware with a typo in it, and the whole `-m gpu` suite doubles as a CPU regression suite for kernel work.
Tests sized for the real machine are skipped (``TOO_LARGE``).  Nothing under wav2letter_pytorch_b200/ imports this module."""
import contextlib
import time
import types

import torch
from torch.overrides import TorchFunctionMode

# `-m gpu` tests that only make sense at full size / on real devices (node-id substrings)
TOO_LARGE = ["test_ctc_full_size_properties", "test_decode_full_size_properties", "test_conv_full_size_layer",
             "test_baseline_config1_forward_ctc_decode", "test_conv_slab_mode_opt_in", "test_ctc_random[False-64-750-225-29]",
             "test_ctc_random[True-64-750-225-29]", "test_peer_gradient_reducer_two_ranks",
             "test_dw_tiled_equals_default[64-751-",
             "test_w2l20_train_step_parity", "test_jasper10x5_block_shapes_parity", "test_conv_full_size_backward_vs_torch",
             "test_ctc_full_size_gradient_vs_torch",
             "test_gpu_zz_graph"]          # CUDA graph capture: a property of the real runtime
# `-m gpu` tests that assert the ABSENCE of a CPU path (here every tensor answers is_cuda = True)
NOT_APPLICABLE = ["test_ctc_module_matches_torch"]


def _is_cuda_dev(d):
    if isinstance(d, torch.device):
        return d.type == "cuda"
    if isinstance(d, str):
        return d.startswith("cuda")
    return False


class _FakeEvent:
    """host clock instead of the device's: elapsed times are those of the emulation (meaningless as measurements, non-zero as numbers)"""

    def __init__(self, *a, **k):
        self.t = time.perf_counter()

    def record(self, *a, **k):
        self.t = time.perf_counter()
        return self

    def synchronize(self):
        pass

    def wait(self, *a, **k):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class _FakeStream:
    cuda_stream = 0
    device = torch.device("cpu")

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def record_event(self, ev=None):
        return ev or _FakeEvent()

    def synchronize(self):
        pass

    def query(self):
        return True


_STREAM = _FakeStream()
_GETTERS = {}


class FakeCudaMode(TorchFunctionMode):
    def __torch_function__(self, func, types, args=(), kwargs=None):
        kwargs = dict(kwargs or {})
        name = getattr(func, "__name__", "")
        if name == "__get__":                              # attribute getters arrive as method-wrappers of the getset descriptor
            if getattr(func, "__self__", None) is _GETTERS["is_cuda"]:
                return True
            return func(*args, **kwargs)
        if name == "cuda" and args and isinstance(args[0], torch.Tensor):
            # a COPY, like the real thing: a raw pointer taken from a `.cuda()` temporary then dangles here too (the CPU allocator hands
            # the freed block to the next temporary), which is how round 1's red GPU test would have been caught on the host
            return args[0].clone()
        if name == "pin_memory" and args and isinstance(args[0], torch.Tensor):
            return args[0]
        if name == "is_pinned":
            return True
        if name == "record_stream":
            return None
        if _is_cuda_dev(kwargs.get("device")):
            kwargs["device"] = "cpu"
        if any(_is_cuda_dev(a) for a in args):
            args = tuple("cpu" if _is_cuda_dev(a) else a for a in args)
        return func(*args, **kwargs)


_enabled = []


def enable():
    """idempotent; stays on for the rest of the process"""
    if _enabled:
        return
    import _emu_cabi
    from wav2letter_pytorch_b200 import _lib
    from wav2letter_pytorch_b200 import functional as F
    _GETTERS["is_cuda"] = torch._C.TensorBase.__dict__["is_cuda"]
    lib = _emu_cabi.library()
    _lib.load = lambda: lib
    # every module of the package sees a `torch` whose empty() is poisoned (an unwritten output / a read of memory nobody wrote
    # shows in the results) and hands out 256-byte aligned byte buffers (as cudaMalloc does and the C wrappers require)
    import importlib
    import pkgutil
    import wav2letter_pytorch_b200 as pkg
    proxy = _emu_cabi._TorchProxy()
    for m in pkgutil.iter_modules(pkg.__path__):
        if m.name.startswith("lib") or m.ispkg:            # the shared library sits in the package directory too
            continue
        mod = importlib.import_module("wav2letter_pytorch_b200." + m.name)
        if getattr(mod, "torch", None) is torch:
            mod.torch = proxy
    F._need_cuda = lambda *ts: None                      # (the function mode is not active inside autograd's backward calls)
    c = torch.cuda
    c.is_available = lambda: True
    c.device_count = lambda: 1
    c.current_device = lambda: 0
    c.set_device = lambda *a, **k: None
    c.synchronize = lambda *a, **k: None
    c.device = lambda *a, **k: contextlib.nullcontext()
    c.stream = lambda *a, **k: contextlib.nullcontext()
    c.current_stream = lambda *a, **k: _STREAM
    c.Stream = _FakeStream
    c.Event = _FakeEvent
    c.mem_get_info = lambda *a, **k: (1 << 20, 2 << 20)
    c.get_device_properties = lambda *a, **k: types.SimpleNamespace(name="emulated sm_100a", multi_processor_count=148, total_memory=180 << 30,
                                                                    major=10, minor=0)
    c.get_device_name = lambda *a, **k: "emulated sm_100a"
    c.memory_stats = lambda *a, **k: {}
    c.empty_cache = lambda: None
    c.manual_seed_all = lambda *a, **k: None
    # the same answers outside the function mode (autograd runs backward() of custom Functions without it)
    torch.Tensor.record_stream = lambda self, stream: None
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.Tensor.is_cuda = property(lambda self: True)
    mode = FakeCudaMode()
    mode.__enter__()
    _enabled.append(mode)
