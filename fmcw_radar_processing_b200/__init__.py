"""B200-native range-Doppler-STFT chain behind the function surface of
alepnabil/fmcw_radar_processing (radar_processing.m / radar_processing_with_azure.m)."""
from ._lib import FmcwError, LAYOUT_FREQ_MAJOR, LAYOUT_TIME_MAJOR  # noqa: F401
from .config import fmcw_configurations  # noqa: F401

__all__ = ["FmcwError", "fmcw_configurations", "LAYOUT_TIME_MAJOR", "LAYOUT_FREQ_MAJOR"]
