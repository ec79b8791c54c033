"""agatha_b200 -- B200-native guided sequence alignment (banded affine-gap extension with Z-drop).

The product is libagatha_b200.so (hand-written sm_100a CUDA behind the C ABI in include/agatha_b200.h); this
package is the thin Python mirror of that ABI used by the tests and bench.py. PyTorch is used only to own
device memory and streams. There is no CPU fallback: importing works anywhere, computing needs the library
and a CUDA device.
"""
from ._lib import Params, DEFAULT_PARAMS, STOP_END, STOP_ZDROP, STOP_BANDEXIT, lib, lib_path, AgathaError  # noqa: F401
from .device_api import (extend_device, pack_device, apply_ops_device, stage_pairs, align_pairs_device, launch_count,  # noqa: F401
                         measure_int_peak)
from .host_api import (align_job, align_job_starts, align_pairs, synth_pairs, bucket_order, shard_pairs, count_cells, fasta_load,  # noqa: F401
                       write_fasta, device_count, Stream, RESULT_DTYPE, stage_batch, pack_batch)
from ._lib import make_params  # noqa: F401
