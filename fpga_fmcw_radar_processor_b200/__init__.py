"""Importable alias of the product package directory `fpga-fmcw-radar-processor_b200/`
(hyphens are not legal in a Python module name)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                                 "fpga-fmcw-radar-processor_b200"))
from .rdp import (DET_DTYPE, TRACK_DTYPE, DeviceBuffer, RadarCore, RdpError, load_library, process_batch,  # noqa: E402,F401
                  process_cpi)
from .io_text import read_detections, read_tracks, write_detections, write_tracks  # noqa: E402,F401
from .sharding import merge_detections, shard_cpis  # noqa: E402,F401
from .sharded import ShardedCpi  # noqa: E402,F401
from .exchange import SocketExchange  # noqa: E402,F401
from .scenario import TacticalScenario, synthetic_cube  # noqa: E402,F401
