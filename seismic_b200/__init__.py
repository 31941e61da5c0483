"""seismic_b200 — B200-native implementation of Seismic's query-time hot path.

`import seismic_b200 as seismic` gives the reference's Python surface (SeismicIndex, SeismicIndexLV,
SeismicIndexRaw, SeismicIndexRawLV, SeismicIndexDotVByte, SeismicDataset, SeismicDatasetLV, get_seismic_string);
Dataset / HostIndex / GpuIndex are the low-level objects over the C ABI (include/seismic_b200.h)."""
from .core import Dataset, GpuGroup, GpuIndex, HostIndex, make_config, pinned_array, recall_at_k  # noqa: F401
from .api import (  # noqa: F401
    SeismicDataset, SeismicDatasetLV, SeismicIndex, SeismicIndexDotVByte, SeismicIndexLV, SeismicIndexRaw,
    SeismicIndexRawLV, get_seismic_string,
)

__version__ = "0.1.0"
