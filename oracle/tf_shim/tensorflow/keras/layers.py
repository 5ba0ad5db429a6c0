from .._impl import (Activation, BatchNormalization, Concatenate, Conv1D, Dense, Dot, Dropout, Embedding,  # noqa: F401
                     Input, InputLayer, Lambda, Layer, Reshape, TimeDistributed, _Unsupported)


def __getattr__(name):  # GRU, Masking, ... referenced by lstur.py / npa.py at call time only
    return type(name, (_Unsupported,), {})
