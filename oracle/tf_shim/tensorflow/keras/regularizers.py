from .._impl import L2  # noqa: F401


def l2(l2=0.01):
    return L2(l2)
