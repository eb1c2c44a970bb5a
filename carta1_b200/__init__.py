"""carta1_b200 -- B200-native ATRAC1 encode/decode hot path of aynik/carta1.

The compute lives in libcarta1_b200.so (hand-written sm_100a CUDA behind the C ABI of
include/carta1_b200.h).  This package is the host-side mirror of the reference's
JavaScript surface (codec/index.js:26-47) used by the parity tests and the benchmark.
"""
from ._lib import (AEA_HEADER, FRAME, SU_BYTES, Carta1Error, Context, EncOpts, StreamDecoder,
                   StreamEncoder, Tables, aea_parse_header, aea_write_header, default_tables,
                   load, make_enc_opts)

__all__ = [
    "AEA_HEADER", "FRAME", "SU_BYTES", "Carta1Error", "Context", "EncOpts", "StreamDecoder",
    "StreamEncoder", "Tables", "aea_parse_header", "aea_write_header", "default_tables", "load",
    "make_enc_opts",
]
