"""B200-native PSD/STI hot path behind PySpectrogram's ``drfProc`` API.

``pyspectrogram_b200.drfProc`` mirrors the reference module (``import drfProc as dp``);
``pyspectrogram_b200.engine`` is the device-resident API; ``pyspectrogram_b200.dist`` shards
columns over the GPUs of one box.  All compute lives in ``libpsgb200.so`` (hand-written sm_100a
CUDA, C ABI in ``include/psg_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"
