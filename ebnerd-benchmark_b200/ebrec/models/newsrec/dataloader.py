"""Batch feed of the newsrec models -- same classes, fields and output contract as the
reference's src/ebrec/models/newsrec/dataloader.py:19-180, 266-419.

Contract kept (checked by tests ported from test/dataloader/test_newsrec.py:66-105):
  train:  ((his [B,H,T] int, pred [B,C,T] int), y [B,C] int)
  eval :  ((his [sumN,H,T], pred [sumN,1,T]), y [sumN,1])     (history repeated per candidate)
  len(loader) == ceil(n_rows / batch_size); the last batch may be short.

Differences, all below the contract:
  * the reference is built on polars + ``tf.keras.utils.Sequence``; neither exists in the B200
    image, so ``behaviors`` may be a polars/pandas DataFrame, a pyarrow Table or a plain
    ``dict`` of equal-length columns -- they are read once into Python lists / numpy;
  * article ids are mapped to lookup-row indices ONCE at construction for every loader (the
    reference does it per batch in NRMSDataLoader and at init in NRMSDataLoaderPretransform,
    dataloader.py:68-81 vs 130-144 -- same result);
  * lookup indices are plain ints (reference: one-element polars Series, hence its squeeze(axis=2)).

Helper semantics follow src/ebrec/utils/_python.py:370-388 (repeat_by_list_values_from_matrix)
and :412-484 (create_lookup_objects: row 0 = unknown = zeros or mean; indices start at 1).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import numpy as np

DEFAULT_INVIEW_ARTICLES_COL = "article_ids_inview"  # src/ebrec/utils/_constants.py
DEFAULT_LABELS_COL = "labels"
DEFAULT_USER_COL = "user_id"


# ----------------------------------------------------------------------------------------------
# helpers (reference: ebrec.utils._python)
# ----------------------------------------------------------------------------------------------
def repeat_by_list_values_from_matrix(input_array, matrix: np.ndarray, repeats) -> np.ndarray:
    """np.repeat(matrix[input_array], repeats, axis=0)   (_python.py:370-388)."""
    return np.repeat(matrix[np.asarray(input_array)], repeats=np.asarray(repeats), axis=0)


def create_lookup_objects(lookup_dictionary: dict, unknown_representation: str):
    """(id -> row index starting at 1, matrix with an extra row 0 for unknown ids)   (_python.py:412-484)."""
    lookup_indexes = {id_: i for i, id_ in enumerate(lookup_dictionary, start=1)}
    lookup_matrix = np.array(list(lookup_dictionary.values()))
    if unknown_representation == "zeros":
        unknown = np.zeros(lookup_matrix.shape[1:], dtype=lookup_matrix.dtype)
    elif unknown_representation == "mean":
        unknown = np.mean(lookup_matrix, axis=0, dtype=lookup_matrix.dtype)
    else:
        raise ValueError(
            f"'{unknown_representation}' is not a specified method. Can be either 'zeros' or 'mean'.")
    lookup_matrix = np.concatenate([unknown[None], lookup_matrix], axis=0)
    return lookup_indexes, lookup_matrix


def _columns(frame) -> dict[str, list]:
    """Read any supported table type into {column: python list}."""
    if isinstance(frame, dict):
        return {k: list(v) for k, v in frame.items()}
    if hasattr(frame, "to_pydict"):  # pyarrow.Table
        return frame.to_pydict()
    if hasattr(frame, "to_dict"):
        try:  # polars
            return frame.to_dict(as_series=False)
        except TypeError:  # pandas
            return {k: list(v) for k, v in frame.to_dict("list").items()}
    raise TypeError(f"unsupported behaviors table type {type(frame)!r}")


def _map_ids(rows: list, mapping: dict, unknown: int) -> list[list[int]]:
    """map_list_article_id_to_value(...fill_nulls=[0]) of _articles_behaviors.py:102-127."""
    get = mapping.get
    return [[get(a, unknown) for a in (row if row is not None else [])] for row in rows]


# ----------------------------------------------------------------------------------------------
@dataclass
class NewsrecDataLoader:
    """A DataLoader for news recommendation (dataloader.py:19-63); a Keras-Sequence-shaped object:
    ``len(loader)`` batches, ``loader[i] -> (inputs_tuple, y)``."""

    behaviors: Any
    history_column: str
    article_dict: dict
    unknown_representation: str
    eval_mode: bool = False
    batch_size: int = 32
    inview_col: str = DEFAULT_INVIEW_ARTICLES_COL
    labels_col: str = DEFAULT_LABELS_COL
    user_col: str = DEFAULT_USER_COL
    kwargs: dict = field(default=None)

    def __post_init__(self):
        self.lookup_article_index, self.lookup_article_matrix = create_lookup_objects(
            self.article_dict, unknown_representation=self.unknown_representation)
        self.unknown_index = [0]
        self.X, self.y = self.load_data()
        if self.kwargs is not None:
            self.set_kwargs(self.kwargs)

    def __len__(self) -> int:
        return int(np.ceil(self._n / float(self.batch_size)))

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __getitem__(self, idx):
        raise ValueError("Function '__getitem__' needs to be implemented.")

    def load_data(self):
        cols = _columns(self.behaviors)
        self._n = len(cols[self.inview_col])
        X = {k: v for k, v in cols.items() if k != self.labels_col}
        X["n_samples"] = [len(r) for r in cols[self.inview_col]]
        y = cols[self.labels_col]
        # article id -> lookup row index, once
        self._hist_idx = _map_ids(cols[self.history_column], self.lookup_article_index, 0)
        self._inview_idx = _map_ids(cols[self.inview_col], self.lookup_article_index, 0)
        return X, y

    def set_kwargs(self, kwargs: dict):
        for key, value in kwargs.items():
            setattr(self, key, value)

    def _slice(self, idx):
        if idx < 0 or idx >= len(self):
            raise IndexError(idx)
        return slice(idx * self.batch_size, (idx + 1) * self.batch_size)


@dataclass
class NRMSDataLoader(NewsrecDataLoader):
    def __getitem__(self, idx):
        """his_input_title [samples, history_size, title_size]; pred_input_title [samples, npratio(+1), title_size];
        batch_y [samples, npratio(+1)]   (dataloader.py:83-119)."""
        sl = self._slice(idx)
        hist, inview, y = self._hist_idx[sl], self._inview_idx[sl], self.y[sl]
        matrix = self.lookup_article_matrix
        if self.eval_mode:
            repeats = np.array([len(r) for r in inview])
            batch_y = np.array([v for row in y for v in row]).reshape(-1, 1)
            his_input_title = repeat_by_list_values_from_matrix(np.array(hist), matrix=matrix, repeats=repeats)
            pred_input_title = matrix[np.array([v for row in inview for v in row], dtype=np.int64)][:, None, :]
        else:
            batch_y = np.array(y)
            his_input_title = matrix[np.array(hist)]
            pred_input_title = matrix[np.array(inview)]
        return (his_input_title, pred_input_title), batch_y


@dataclass
class NRMSDataLoaderDevice(NRMSDataLoader):
    """Device-resident batch feed (not in the reference): the
    ``[n_articles + 1, title_size]`` token matrix (``lookup_article_matrix``) is uploaded to HBM once by the model
    (``fit`` / ``predict`` see ``device_feed``) and a batch is only the article ROW INDICES --
    train: ((his_idx [B,H] int32, pred_idx [B,C] int32), y [B,C]); eval: ((his_idx [sumN,H], pred_idx [sumN,1]), y [sumN,1])
    -- 30x fewer host->device bytes than the token tensors of NRMSDataLoader and no per-batch host gather
    (dataloader.py:110-118).  ``lookup_article_matrix[batch indices]`` reproduces NRMSDataLoader's batch exactly."""

    device_feed = True

    def __getitem__(self, idx):
        sl = self._slice(idx)
        hist, inview, y = self._hist_idx[sl], self._inview_idx[sl], self.y[sl]
        if self.eval_mode:
            repeats = np.array([len(r) for r in inview])
            batch_y = np.array([v for row in y for v in row]).reshape(-1, 1)
            his_idx = np.repeat(np.array(hist, dtype=np.int32), repeats, axis=0)
            pred_idx = np.array([v for row in inview for v in row], dtype=np.int32)[:, None]
        else:
            batch_y = np.array(y)
            his_idx = np.array(hist, dtype=np.int32)
            pred_idx = np.array(inview, dtype=np.int32)
        return (his_idx, pred_idx), batch_y


@dataclass
class NRMSDocVecDataLoaderDevice(NRMSDataLoaderDevice):
    """Device-resident feed for NRMSDocVec: ``article_dict`` maps article ids
    to 768-d document vectors, so ``lookup_article_matrix`` is the float [n_articles + 1, 768] matrix (386 MB for
    ebnerd_large) -- uploaded to HBM once; batches are the same row-index tuples as NRMSDataLoaderDevice."""


@dataclass
class NRMSDataLoaderPretransform(NRMSDataLoader):
    """Reference: pre-transforms the whole frame in __post_init__ (dataloader.py:122-180).  Every loader
    of this build already does that, so this is the same class under the reference's name."""


@dataclass(kw_only=True)
class NAMLDataLoader(NewsrecDataLoader):
    """Eval mode not implemented (dataloader.py:266-419).  Returns the 8-tuple
    (his_title, his_body, his_vert, his_subvert, pred_title, pred_body, pred_vert, pred_subvert), y."""

    unknown_category_value: int = 0
    unknown_subcategory_value: int = 0
    body_mapping: dict = None
    category_mapping: dict = None
    subcategory_mapping: dict = None

    def __post_init__(self):
        self.lookup_article_index_body, self.lookup_article_matrix_body = create_lookup_objects(
            self.body_mapping, unknown_representation=self.unknown_representation)
        if self.eval_mode:
            raise ValueError("'eval_mode = True' is not implemented for NAML")
        super().__post_init__()
        cols = _columns(self.behaviors)
        hist, inview = cols[self.history_column], cols[self.inview_col]
        self._hist_body = _map_ids(hist, self.lookup_article_index_body, 0)
        self._inview_body = _map_ids(inview, self.lookup_article_index_body, 0)
        self._hist_cat = _map_ids(hist, self.category_mapping, self.unknown_category_value)
        self._inview_cat = _map_ids(inview, self.category_mapping, self.unknown_category_value)
        self._hist_sub = _map_ids(hist, self.subcategory_mapping, self.unknown_subcategory_value)
        self._inview_sub = _map_ids(inview, self.subcategory_mapping, self.unknown_subcategory_value)

    def __getitem__(self, idx):
        sl = self._slice(idx)
        batch_y = np.array(self.y[sl])
        mt, mb = self.lookup_article_matrix, self.lookup_article_matrix_body
        return (
            mt[np.array(self._hist_idx[sl])],
            mb[np.array(self._hist_body[sl])],
            np.array(self._hist_cat[sl])[:, :, np.newaxis],
            np.array(self._hist_sub[sl])[:, :, np.newaxis],
            mt[np.array(self._inview_idx[sl])],
            mb[np.array(self._inview_body[sl])],
            np.array(self._inview_cat[sl])[:, :, np.newaxis],
            np.array(self._inview_sub[sl])[:, :, np.newaxis],
        ), batch_y


@dataclass(kw_only=True)
class NAMLDataLoaderDevice(NAMLDataLoader):
    """Device-resident feed for NAML: title / body token matrices live in HBM (uploaded once by the model), a batch
    carries article row indices for the two text views and the category ids:
    (his_title_idx [B,H], his_body_idx [B,H], his_vert [B,H,1], his_subvert [B,H,1], pred_title_idx [B,C], ...), y.
    ``lookup_article_matrix[idx]`` / ``lookup_article_matrix_body[idx]`` reproduce NAMLDataLoader's batch exactly."""

    device_feed = True

    def __getitem__(self, idx):
        sl = self._slice(idx)
        batch_y = np.array(self.y[sl])
        i32 = np.int32
        return (
            np.array(self._hist_idx[sl], dtype=i32),
            np.array(self._hist_body[sl], dtype=i32),
            np.array(self._hist_cat[sl])[:, :, np.newaxis],
            np.array(self._hist_sub[sl])[:, :, np.newaxis],
            np.array(self._inview_idx[sl], dtype=i32),
            np.array(self._inview_body[sl], dtype=i32),
            np.array(self._inview_cat[sl])[:, :, np.newaxis],
            np.array(self._inview_sub[sl])[:, :, np.newaxis],
        ), batch_y
