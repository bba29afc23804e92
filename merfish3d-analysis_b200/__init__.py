"""B200-native per-pixel barcode decoding for merfish3d-analysis (PixelDecoder hot path).

Import as ``merfish3d_analysis_b200``.  Heavy submodules (the ctypes binding to
``libm3d_b200.so`` and the torch host code) are imported lazily so that pure-host helpers
(``synthetic``, ``codebook``) work without a GPU.
"""

__version__ = "0.1.0"
