from .._impl import Input, Model  # noqa: F401
from . import backend, initializers, layers, optimizers, regularizers, utils  # noqa: F401
