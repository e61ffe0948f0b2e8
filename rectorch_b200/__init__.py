"""rectorch_b200 -- a Blackwell-native (sm_100a) engine behind the MultiVAE / MultiDAE path of
makgyver/rectorch: ``rectorch_b200.nets.MultiVAE_net/MultiDAE_net``,
``rectorch_b200.models.MultiVAE/MultiDAE``, ``rectorch_b200.samplers.DataSampler``,
``rectorch_b200.evaluation.evaluate/ValidFunc`` and ``rectorch_b200.metrics.Metrics`` keep the
reference's signatures; all arithmetic runs in hand-written CUDA kernels (libb200vae.so, C ABI in
include/b200vae.h).  Importing the package does not need a GPU; computing anything does.
"""
__version__ = "0.1.0"
__all__ = ["nets", "models", "samplers", "evaluation", "metrics", "synth"]
