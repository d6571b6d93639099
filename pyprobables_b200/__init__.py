"""pyprobables_b200 -- B200-native batch engine for pyprobables' hash-then-scatter hot path.

Same class names, constructor arguments, properties, exceptions and `hash_function` plugin seam as
`probables` (barrust/pyprobables v0.7.0) for BloomFilter, CountingBloomFilter, Expanding/RotatingBloomFilter, CountMinSketch (+Mean, +MeanMin)
CuckooFilter and CountingCuckooFilter, with the batch methods `add_many` / `check_many` added.  State lives in GPU memory; every
add/check runs in hand-written sm_100a CUDA kernels behind the C ABI of include/pb200.h.
There is no CPU fallback: without libpb200.so and a CUDA device the constructors raise.
"""

from . import hashes
from ._native import Context, NativeError, NoDeviceError, build, default_context, device_count
from .bloom import BloomFilter
from .countingbloom import CountingBloomFilter
from .countingcuckoo import CountingCuckooFilter
from .countminsketch import CountMeanMinSketch, CountMeanSketch, CountMinSketch, HeavyHitters, StreamThreshold
from .cuckoo import CuckooFilter
from .expandingbloom import ExpandingBloomFilter, RotatingBloomFilter
from .exceptions import (
    CountMinSketchError,
    CuckooFilterFullError,
    InitializationError,
    NotSupportedError,
    RotatingBloomFilterError,
    ProbablesBaseException,
    SimilarityError,
)
from .keys import KeyBatch, pack_keys

__version__ = "0.1.0"

__all__ = [
    "BloomFilter",
    "CountingBloomFilter",
    "ExpandingBloomFilter",
    "RotatingBloomFilter",
    "RotatingBloomFilterError",
    "CountMinSketch",
    "CountMeanSketch",
    "CountMeanMinSketch",
    "HeavyHitters",
    "StreamThreshold",
    "CuckooFilter",
    "CountingCuckooFilter",
    "InitializationError",
    "NotSupportedError",
    "ProbablesBaseException",
    "CuckooFilterFullError",
    "CountMinSketchError",
    "SimilarityError",
    "Context",
    "default_context",
    "device_count",
    "pack_keys",
    "KeyBatch",
    "hashes",
    "build",
    "NativeError",
    "NoDeviceError",
]
