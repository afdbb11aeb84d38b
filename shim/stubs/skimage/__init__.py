"""Stand-in for scikit-image (absent from this image): only skimage.filters.threshold_multiotsu is used by the
reference, for the load-time noise statistics of data/NeuronDataset.py:90-93."""
from . import filters  # noqa: F401
