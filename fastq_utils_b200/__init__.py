"""fastq_utils_b200 — B200-native fastq_info hot path (libfastq_gpu) and its Python host-side mirror.

The product is the C-ABI library `libfastq_gpu.so` (include/fastq_gpu.h); this package only binds it with ctypes and
adds the multi-GPU orchestration (`dist`) that uses torch.distributed for the plumbing.  There is no CPU fallback:
importing works anywhere, creating a context needs a CUDA device.
"""
from .api import (FastqInfo, fastq_info, reader_tool, trim_poly_at, filterpair, lib, MODE_READER, MODE_SINGLE, MODE_INDEX, MODE_INDEX_PAIR, MODE_INTERLEAVED,  # noqa: F401
                  MODE_SORTED_PAIR, KERNEL_CLASSES, FLAG_PAIRED_NAMES, FLAG_EXTERNAL_INDEX, FLAG_TWO_PASS, FLAG_BORROW_FOR_CALL,
                  synth_illumina, synth_longreads, illumina_record_bytes)

__all__ = ["FastqInfo", "fastq_info", "reader_tool", "trim_poly_at", "filterpair", "lib", "synth_illumina", "synth_longreads", "illumina_record_bytes"]
