"""seismic_b200 — B200-native implementation of Seismic's query-time hot path."""
from .core import Dataset, GpuIndex, HostIndex, make_config, recall_at_k  # noqa: F401

__version__ = "0.1.0"
