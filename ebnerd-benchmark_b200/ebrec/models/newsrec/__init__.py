"""B200-native drop-in for ``ebrec.models.newsrec``.

Reference: src/ebrec/models/newsrec/__init__.py:1-4 exports NPAModel, LSTURModel,
NRMSModel, NAMLModel.  This build covers the NRMS family named by the north star
(NRMS / NRMSDocVec / NAML); LSTUR and NPA are out of scope.
"""
from .naml import NAMLModel  # noqa: F401
from .nrms import NRMSModel  # noqa: F401
