"""gpuar_b200 -- B200-native (sm_100a) GPUAR codec: hand-written CUDA kernels behind the
C ABI of include/gpuar_b200.h, plus the thin host plumbing around it.

    gpuar_b200._lib     ctypes binding of libgpuar_b200.so (fails loudly if not built)
    gpuar_b200.codec    device-resident encode / index / decode and the host-buffer pipeline
    gpuar_b200.shard    packet-range sharding across the GPUs of one box
    gpuar_b200.datagen  synthetic inputs
    gpuar_b200/csrc     the kernels, the C ABI and the C++ host side (gpuar CLI)
"""
__version__ = "0.1.0"
