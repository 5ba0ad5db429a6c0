from .._impl import backend as _b

globals().update({k: getattr(_b, k) for k in dir(_b) if not k.startswith("_")})
