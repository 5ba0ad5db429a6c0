"""`tensorflow` stand-in (oracle/tf_shim): see _impl.py.  TEST INFRASTRUCTURE ONLY -- never on the product path."""
from . import keras  # noqa: F401
from ._impl import matmul  # noqa: F401

__version__ = "0.0-ebk-shim"
IS_EBK_SHIM = True


class random:  # tf.random.set_seed (nrms.py:36, base_model.py:36)
    @staticmethod
    def set_seed(seed):
        return None
