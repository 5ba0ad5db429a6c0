from .._impl import GlorotUniform, Ones, RandomUniform, Zeros  # noqa: F401

glorot_uniform = GlorotUniform
