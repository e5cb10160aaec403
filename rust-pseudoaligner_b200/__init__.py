"""B200-native pseudoalignment hot path of 10XGenomics/rust-pseudoaligner.

The product is `libpsa_b200.so` (C ABI in include/psa.h, CUDA kernels in csrc/).  This
package is the host-side mirror of the crate's API for that path, over ctypes:

    Pseudoaligner.map_read      <- ref src/pseudoaligner.rs:381
    process_reads               <- ref src/pseudoaligner.rs:420-514

There is no CPU implementation here: importing works without a GPU (so that the symbol
table can be checked), every compute call needs the CUDA library and a device.
"""
from .pseudoaligner import (  # noqa: F401
    EQ_NONE, FLAG_ALIGNED, FLAG_MAPPED, HIT_DTYPE, Index, Mapper, Pseudoaligner, PsaError,
    Comm, format_read_data, lib, lib_path, process_reads, process_reads_file, DeviceBatch,
)
