"""wav2letter_pytorch_b200 -- B200-native (sm_100a) hot path of assafmu/wav2letter_pytorch behind the reference's
own Python API: ``Wav2Letter`` / ``Jasper`` modules, ``CTCLoss``, ``GreedyDecoder``, ``Novograd``, Hydra-layout configs.
Every operator calls hand-written CUDA through the C ABI in ``include/w2l_sm100.h``; there is no CPU fallback."""
from . import config, label_sets  # noqa: F401

name_to_model = {}


def _register():
    from .wav2letter import Wav2Letter
    name_to_model["wav2letter"] = Wav2Letter
    try:
        from .jasper import Jasper
        name_to_model["jasper"] = Jasper
    except ImportError:
        pass


_register()


def reserve_device_memory(device=None, gib=64, fraction=0.6):
    """Pre-size PyTorch's caching allocator with ONE large segment (then released to the cache): every later activation /
    workspace allocation of a training step is carved out of it, so steady-state steps never call cudaMalloc -- which
    synchronises the device and, with the host running several steps ahead of a 180 GB GPU, otherwise keeps happening long
    after warm-up (measured: 68 cudaMalloc calls and a 52 ms outlier inside 20 timed 39 ms steps)."""
    import torch
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    free, _total = torch.cuda.mem_get_info(device)
    n = min(int(gib) << 30, int(free * fraction))
    if n > 0:
        try:
            block = torch.empty(n, dtype=torch.uint8, device=device)
            del block
        except RuntimeError:                      # not enough contiguous memory: run without the reservation
            return 0
    return n
