"""b200-stereo: B200-native (sm_100a) drop-in for gpuimageproc's stereo hot path.

The compute lives in libb200stereo.so (CUDA kernels + C ABI, see include/b200_stereo.h); this package is the
Python host-side mirror of gpuimageproc::GpuStereoProcessor.  There is no CPU fallback.
"""
from . import _capi  # noqa: F401
from .processor import *  # noqa: F401,F403
from .processor import GpuStereoProcessor, GpuStereoPool  # noqa: F401
