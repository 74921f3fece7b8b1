"""pywfo_b200 -- B200-native CI wavefunction-overlap hot path.

Drop-in for the path ``pywfo.main.overlaps*`` / ``pywfo.main.moovlp*`` /
``pywfo.dets.wfoverlap`` of eljost/pywfo: same Python signatures, the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of
``include/wfo_b200.h``.  There is no CPU fallback.
"""
from . import excitations  # noqa: F401

__all__ = ["main", "dets", "helpers", "excitations", "dist", "trajectory", "formats"]
__version__ = "0.1.0"
