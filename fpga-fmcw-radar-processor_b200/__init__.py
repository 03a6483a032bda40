"""B200-native range-Doppler + OS-CFAR path (drop-in for the `radar_core` chain of
Aurellia-Beam/fpga-fmcw-radar-processor).  Import as `fpga_fmcw_radar_processor_b200`
(the directory name carries the reference's hyphens, which Python cannot import directly)."""
from .rdp import (DET_DTYPE, TRACK_DTYPE, DeviceBuffer, RadarCore, RdpError, load_library, process_batch,  # noqa: F401
                  process_cpi)
from .io_text import read_detections, read_tracks, write_detections, write_tracks  # noqa: F401
from .sharding import merge_detections, shard_cpis  # noqa: F401
from .sharded import ShardedCpi  # noqa: E402,F401
from .exchange import SocketExchange  # noqa: E402,F401
