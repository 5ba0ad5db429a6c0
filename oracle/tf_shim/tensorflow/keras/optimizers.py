from .._impl import Adam  # noqa: F401
