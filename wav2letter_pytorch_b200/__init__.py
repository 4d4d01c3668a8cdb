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
