"""Host data preparation for the newsrec dataloaders without polars.

The reference prepares its behaviour frames with polars (`src/ebrec/utils/_behaviors.py`); polars is not
installable in the B200 image, while the dataloaders of this package accept a plain ``dict`` of equal-length
columns (or a pyarrow Table).  This module restates the four steps the training scripts run before they build a
dataloader (examples/reproducibility_scripts/ebnerd_nrms.py:158-200) on such "frames":

    ebnerd_from_path            _behaviors.py:161-192   behaviors.parquet LEFT JOIN truncated/padded history.parquet
    truncate_history            _behaviors.py:582-654   tail(history_size) of a list column, optional LEFT padding
    sampling_strategy_wu2019    _behaviors.py:423-579   one row per clicked article: npratio sampled negatives + the click
    create_binary_labels_column _behaviors.py:22-107    labels[i] = inview[i] in clicked

A frame is ``{column name: list}``; every function returns a NEW frame with the same columns (plus the one it
adds) and never mutates its input.  Random choices use ``numpy.random.default_rng(seed)`` -- polars' own sampler
cannot be reproduced, so seeded outputs differ from the reference's docstring tables in WHICH negatives are drawn,
not in structure (counts, click last before shuffling, ``None`` negatives when an impression has none).
"""
from __future__ import annotations

from pathlib import Path
from typing import Any

import numpy as np

DEFAULT_USER_COL = "user_id"                          # src/ebrec/utils/_constants.py
DEFAULT_HISTORY_ARTICLE_ID_COL = "article_id_fixed"
DEFAULT_INVIEW_ARTICLES_COL = "article_ids_inview"
DEFAULT_CLICKED_ARTICLES_COL = "article_ids_clicked"
DEFAULT_LABELS_COL = "labels"

Frame = dict


def _as_frame(table) -> Frame:
    if isinstance(table, dict):
        return {k: list(v) for k, v in table.items()}
    if hasattr(table, "to_pydict"):          # pyarrow.Table
        return table.to_pydict()
    raise TypeError(f"unsupported frame type {type(table)!r} (dict of columns or pyarrow.Table)")


def _n_rows(frame: Frame) -> int:
    return len(next(iter(frame.values()))) if frame else 0


def _need(frame: Frame, columns) -> None:
    missing = [c for c in columns if c not in frame]
    if missing:
        raise ValueError(f"Invalid input provided. The dataframe does not contain columns {missing}.")


def truncate_history(frame, column: str, history_size: int, padding_value: Any = None) -> Frame:
    """Keep the LAST ``history_size`` items of every list in ``column`` (histories are in ascending time order);
    with ``padding_value`` shorter lists are padded on the LEFT to exactly ``history_size``."""
    f = _as_frame(frame)
    _need(f, [column])
    out = []
    for row in f[column]:
        row = list(row) if row is not None else []
        row = row[-history_size:] if history_size > 0 else []
        if padding_value is not None and len(row) < history_size:
            row = [padding_value] * (history_size - len(row)) + row
        out.append(row)
    f[column] = out
    return f


def ebnerd_from_path(path, history_size: int = 30, padding: int = 0, user_col: str = DEFAULT_USER_COL,
                     history_aids_col: str = DEFAULT_HISTORY_ARTICLE_ID_COL) -> Frame:
    """behaviors.parquet with each user's truncated / left-padded click history attached (left join on the user;
    users without a history row get ``None``)."""
    import pyarrow.parquet as pq

    path = Path(path)
    hist = pq.read_table(path / "history.parquet", columns=[user_col, history_aids_col]).to_pydict()
    hist = truncate_history(hist, history_aids_col, history_size, padding_value=padding)
    by_user = dict(zip(hist[user_col], hist[history_aids_col]))
    beh = pq.read_table(path / "behaviors.parquet").to_pydict()
    beh[history_aids_col] = [by_user.get(u) for u in beh[user_col]]
    return beh


def create_binary_labels_column(frame, shuffle: bool = False, seed: int = None,
                                clicked_col: str = DEFAULT_CLICKED_ARTICLES_COL,
                                inview_col: str = DEFAULT_INVIEW_ARTICLES_COL,
                                label_col: str = DEFAULT_LABELS_COL) -> Frame:
    """Adds ``label_col``: for every in-view article 1 if it is among the row's clicked articles else 0 (a missing
    clicked list gives all zeros).  ``shuffle`` permutes each in-view list first (labels follow the new order)."""
    f = _as_frame(frame)
    _need(f, [inview_col, clicked_col])
    rng = np.random.default_rng(seed)
    inview_out, labels = [], []
    for inview, clicked in zip(f[inview_col], f[clicked_col]):
        inview = list(inview) if inview is not None else []
        if shuffle and len(inview) > 1:
            inview = [inview[i] for i in rng.permutation(len(inview))]
        hit = set(clicked) if clicked is not None else set()
        inview_out.append(inview)
        labels.append([1 if a in hit else 0 for a in inview])
    f[inview_col] = inview_out
    f[label_col] = labels
    return f


def sampling_strategy_wu2019(frame, npratio: int, shuffle: bool = False, with_replacement: bool = True, seed: int = None,
                             inview_col: str = DEFAULT_INVIEW_ARTICLES_COL,
                             clicked_col: str = DEFAULT_CLICKED_ARTICLES_COL) -> Frame:
    """Negative sampling of Wu et al. (NPA, KDD'19) as the reference applies it:
    1. the clicked articles are removed from the in-view list (the negatives of the impression);
    2. the row is repeated once per clicked article;
    3. ``npratio`` negatives are drawn for each repeated row (with or without replacement; an impression without
       negatives yields ``None`` entries, as the reference's null-filled lists do);
    4. the clicked article is appended LAST, and ``clicked_col`` becomes the one-element list holding it;
    5. ``shuffle`` permutes every resulting in-view list.
    All other columns are repeated unchanged.  Raises ValueError for ``npratio < 0`` and, without replacement,
    when an impression has fewer negatives than ``npratio``."""
    if npratio < 0:
        raise ValueError(f"npratio must be >= 0, got {npratio}")
    f = _as_frame(frame)
    _need(f, [inview_col, clicked_col])
    rng = np.random.default_rng(seed)
    out = {k: [] for k in f}
    n = _n_rows(f)
    for i in range(n):
        clicked = f[clicked_col][i]
        clicked = list(clicked) if clicked is not None else []
        hit = set(clicked)
        negatives = [a for a in (f[inview_col][i] or []) if a not in hit]
        for pos in clicked:
            if not negatives:
                sample = [None] * npratio
            elif with_replacement:
                sample = [negatives[j] for j in rng.integers(0, len(negatives), npratio)]
            else:
                if npratio > len(negatives):
                    raise ValueError("cannot take a larger sample than the total population when with_replacement=False")
                sample = [negatives[j] for j in rng.permutation(len(negatives))[:npratio]]
            row = sample + [pos]
            if shuffle and len(row) > 1:
                row = [row[j] for j in rng.permutation(len(row))]
            for k in f:
                if k == inview_col:
                    out[k].append(row)
                elif k == clicked_col:
                    out[k].append([pos])
                else:
                    out[k].append(f[k][i])
    return out


def select_rows(frame, mask) -> Frame:
    """Rows of ``frame`` where ``mask`` is true (the scripts' ``.filter(...)`` / ``.sample(...)`` steps)."""
    f = _as_frame(frame)
    keep = [i for i, m in enumerate(mask) if m]
    return {k: [v[i] for i in keep] for k, v in f.items()}
