"""darwin_b200 -- B200-native GACT alignment extension (drop-in for the Processor/extender seam of
yatisht/darwin).  CUDA kernels + C-ABI live in csrc/ (built into libdarwin_gact.so); `gact.Processor`
is the Python mirror of the reference's call surface."""
from . import abi, synth  # noqa: F401
from .gact import Processor, DarwinGpuError, load_library, read_params_cfg, scoring_from_cfg, cigar, sam_select  # noqa: F401

__all__ = ["abi", "synth", "Processor", "DarwinGpuError", "load_library", "read_params_cfg", "scoring_from_cfg", "cigar", "sam_select"]
