"""vtaco_b200 — B200-native (sm_100a) implementation of jeffsonyu/VTacO's
convolutional-occupancy hot path behind the reference's module API.

    from vtaco_b200.encoder import encoder_dict           # src/encoder/__init__.py
    from vtaco_b200.conv_onet.models import decoder_dict, ConvolutionalOccupancyNetwork
    from vtaco_b200.conv_onet.generation import Generator3D

CUDA-only: the kernels live in vtaco_b200/lib/libvtaco_b200.so (C ABI in
include/vtaco_b200.h); there is no PyTorch or CPU fallback.
"""
__version__ = '0.1.0'
