"""Entry point named like the reference's server function file (radar_processing_with_azure.m)."""
from .radar_processing import main  # noqa: F401
