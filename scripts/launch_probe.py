"""Fixed cost of a kernel launch as a function of its configuration (empty kernel, CUDA-graph replay of 200
back-to-back launches): plain grid vs thread-block clusters vs the 225 KB dynamic shared memory of the tcgen05 kernels.

    python scripts/launch_probe.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import _lib  # noqa: E402
from rectorch_b200._lib import check  # noqa: E402

torch.cuda.init()
side = torch.cuda.Stream()
REP = 200
for grid, threads, smem, cluster in [(148, 128, 0, 1), (148, 576, 0, 1), (148, 576, 0, 2), (148, 576, 100 * 1024, 1),
                                     (148, 576, 228608, 1), (148, 576, 228608, 2), (148, 320, 228608, 2), (16, 576, 228608, 2),
                                     (1184, 256, 0, 1)]:
    with torch.cuda.stream(side):
        check(_lib.lib().b200vae_probe_launch(grid, threads, smem, cluster, ctypes.c_void_p(side.cuda_stream)))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(REP):
            check(_lib.lib().b200vae_probe_launch(grid, threads, smem, cluster, sp))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("grid %5d x %3d threads, %6d B dynamic smem, cluster %d: %6.2f us per launch" % (
        grid, threads, smem, cluster, e0.elapsed_time(e1) / (5 * REP) * 1e3))
