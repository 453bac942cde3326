"""bdx-b200: B200-native barcode assignment behind BioDemuX.jl's worker boundary.

The directory name carries a dot, so import it through the repo-root shim
``bdx_b200`` (``import bdx_b200 as bdx``).
"""
from .ranges import DynamicRange, parse_dynamic_range, parse_part, resolve  # noqa: F401
from .config import DemuxConfig, build_config  # noqa: F401
from .fileio import preprocess_bc_file, fastq_records  # noqa: F401
from .demux import (execute_demultiplexing, run_pipeline, output_filename, pack_reads,  # noqa: F401
                    RESULT_DTYPE, DETAIL_DTYPE, MATCH, UNKNOWN, AMBIGUOUS)
