import numpy as np


def threshold_multiotsu(image, classes=4, nbins=256):
    """Fixed thresholds (the same ones tests/golden/gen_data_golden.py uses): the reference only needs a cut between
    'noise' and 'signal' pixels to estimate the noise mean / std it pads the canvases with at load time."""
    return np.array([40.0, 90.0, 160.0])[:classes - 1]


def threshold_otsu(image, nbins=256):
    return 90.0
