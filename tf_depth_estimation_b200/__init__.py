"""B200-native view-synthesis loss: the one data-parallel hot path of wrlife/tf_depth_estimation
(utils.py / utils_lr.py / my_losses.py geometry, sampler and loss terms), as hand-written sm_100a CUDA
kernels behind a C ABI (include/vsl.h), with the reference's Python call signatures kept on top.

    from tf_depth_estimation_b200 import ops          # torch-tensor entry points + fused loss
    from tf_depth_estimation_b200.compat import utils_lr, utils, my_losses   # reference-named modules

Importing the package is cheap and works without a GPU; the first op call dlopens libvsl.so and raises if it
is missing -- there is no CPU or eager fallback.
"""
__version__ = '0.1.0'
