"""areafunctionsynthesis_b200 -- B200-native area-function-to-speech hot path.

Thin ctypes layer over the C ABI (include/afs_b200.h, libafs_b200.so).  The
product is the CUDA library; Python is only test/bench plumbing.  There is no
CPU fallback: creating a `Context` without the built library or without a CUDA
device raises.
"""
from .api import (  # noqa: F401
    FRAME_DTYPE,
    TL_ARITHMETIC_FAST,
    TL_ARITHMETIC_REFERENCE,
    AfsError,
    Context,
    Group,
    TdsOptions,
    TdsProbes,
    TdsStream,
    TlOptions,
    area_function_to_frames,
    audio_to_int16,
    chebyshev_lowpass,
    default_frames,
    lf_glottal_flow,
    library_path,
    load_library,
    pinned_empty,
    pinned_free,
    tds_options_default,
    target_sequence_trajectory,
    tl_options_default,
    write_wav_int16,
)

__version__ = "0.2.0"
