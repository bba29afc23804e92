"""B200-native per-pixel barcode decoding for merfish3d-analysis (PixelDecoder hot path).

Import as ``merfish3d_analysis_b200``.  Heavy submodules (the ctypes binding to
``libm3d_b200.so`` and the torch host code) are imported lazily so that pure-host helpers
(``synthetic``, ``codebook``) work without a GPU.
"""

__version__ = "0.1.0"

import os as _os

# The image store decodes up to 48 chunks side by side, each on its own stream (csrc/zarrio.cu).  With the driver's default
# of 8 hardware work queues, streams that share a queue wait for each other's kernels (measured on the B200: the device
# zstd decoder 6.2 -> 22.7 GB/s, the LZ4 store 62.9 -> 79.5 GB/s decoded with 32 queues).  The variable is read when the
# CUDA context is created, so it has to be set before the first CUDA call of the process; an explicit setting wins.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
