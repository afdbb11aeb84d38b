"""Drop-in for the reference's `models` module (train.py:107-108 `from models import *`, eval.py:5): the six names
of the reference's `__all__` (models.py:10-12), served by neuron_gan_b200."""
from neuron_gan_b200.models import *  # noqa: F401,F403
from neuron_gan_b200.models import __all__  # noqa: F401
