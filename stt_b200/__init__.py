"""stt_b200 -- B200-native drop-in for Coqui STT's streaming-inference hot path.

Python surface mirrors the reference binding ``native_client/python/__init__.py:26-380`` (``Model``, ``Stream``,
``Metadata`` ...) over ctypes instead of SWIG, on top of ``libstt_b200.so`` (include/stt_capi.h).
There is no CPU path: loading the library or creating a model without a CUDA device raises.
"""
from .api import (  # noqa: F401
    Model,
    Stream,
    Batch,
    BatchPipeline,
    CandidateTranscript,
    Metadata,
    TokenMetadata,
    lib,
    lib_path,
    version,
    STTError,
)
