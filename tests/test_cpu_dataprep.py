"""Polars-free host data prep (ebrec.models.newsrec.dataprep) against the known answers quoted in the reference's
docstrings (src/ebrec/utils/_behaviors.py:40-75, 493-555, 606-640) and against the dataloaders downstream."""
from pathlib import Path

import numpy as np
import pytest

from ebrec.models.newsrec import dataprep as DP

INV, CLK, LAB = DP.DEFAULT_INVIEW_ARTICLES_COL, DP.DEFAULT_CLICKED_ARTICLES_COL, DP.DEFAULT_LABELS_COL


def test_create_binary_labels_docstring_example():
    df = {INV: [[1, 2, 3], [4, 5, 6], [7, 8]], CLK: [[2, 3, 4], [3, 5], None]}       # _behaviors.py:45-51
    out = DP.create_binary_labels_column(df)
    assert out[LAB] == [[0, 1, 1], [0, 1, 0], [0, 0]]                               # _behaviors.py:52-62, 76-80
    assert out[INV] == df[INV] and df.get(LAB) is None                               # input untouched
    sh = DP.create_binary_labels_column(df, shuffle=True, seed=123)
    assert [sum(r) for r in sh[LAB]] == [2, 1, 0]                                    # _behaviors.py:81-85
    for inv, lab, clk in zip(sh[INV], sh[LAB], df[CLK]):                             # labels follow the shuffled order
        assert lab == [1 if a in (clk or []) else 0 for a in inv]
    assert sorted(sh[INV][0]) == [1, 2, 3]


def test_truncate_history_docstring_example():
    df = {"id": [1, 2, 3], "history": [["a", "b", "c"], ["d", "e", "f", "g"], ["h", "i"]]}   # _behaviors.py:606-608
    assert DP.truncate_history(df, "history", 3)["history"] == [["a", "b", "c"], ["e", "f", "g"], ["h", "i"]]
    assert DP.truncate_history(df, "history", 3, "-")["history"] == [["a", "b", "c"], ["e", "f", "g"], ["-", "h", "i"]]
    assert df["history"][1] == ["d", "e", "f", "g"]


def test_sampling_strategy_wu2019_structure():
    df = {"impression_id": [0, 1, 2, 3], "user_id": [1, 1, 2, 3],                    # _behaviors.py:480-491
          INV: [[1, 2, 3], [1, 2, 3, 4], [1, 2, 3], [1]], CLK: [[1, 2], [1, 3], [1], [1]]}
    for npratio in (1, 2):
        out = DP.sampling_strategy_wu2019(df, npratio=npratio, shuffle=False, with_replacement=True, seed=123)
        # one row per clicked article, other columns repeated                         _behaviors.py:493-506
        assert out["impression_id"] == [0, 0, 1, 1, 2, 3] and out["user_id"] == [1, 1, 1, 1, 2, 3]
        assert out[CLK] == [[1], [2], [1], [3], [1], [1]]
        pools = [{3}, {3}, {2, 4}, {2, 4}, {2, 3}, set()]
        for row, clk, pool in zip(out[INV], out[CLK], pools):
            assert len(row) == npratio + 1 and row[-1] == clk[0]                     # the click comes last
            assert all((a in pool) if pool else (a is None) for a in row[:-1])       # [null, 1] for impression 3
    sh = DP.sampling_strategy_wu2019(df, npratio=2, shuffle=True, seed=5)
    assert all(clk[0] in row and len(row) == 3 for row, clk in zip(sh[INV], sh[CLK]))
    with pytest.raises(ValueError):                                                  # _behaviors.py:535-536
        DP.sampling_strategy_wu2019(df, npratio=2, with_replacement=False, seed=1)
    ok = DP.select_rows(df, [len(r) > 3 for r in df[INV]])
    out = DP.sampling_strategy_wu2019(ok, npratio=2, with_replacement=False, seed=123)
    assert [sorted(r[:-1]) for r in out[INV]] == [[2, 4], [2, 4]] and [r[-1] for r in out[INV]] == [1, 3]
    with pytest.raises(ValueError):
        DP.sampling_strategy_wu2019(df, npratio=-1)


REF = Path("/root/reference/test/data/ebnerd")


@pytest.mark.skipif(not REF.exists(), reason="reference parquet fixtures only exist in the build container")
def test_ebnerd_from_path_to_dataloader_on_reference_fixtures():
    """The script pipeline of ebnerd_nrms.py:158-200 end to end on the reference's own parquet fixtures:
    load + history join -> wu2019 sampling -> labels -> NRMSDataLoader batches of the documented shape."""
    import pyarrow.parquet as pq

    from ebrec.models.newsrec.dataloader import NRMSDataLoader

    H, NP = 7, 4
    df = DP.ebnerd_from_path(REF, history_size=H, padding=0)
    n = pq.read_table(REF / "behaviors.parquet").num_rows
    assert DP._n_rows(df) == n and all(len(h) == H for h in df[DP.DEFAULT_HISTORY_ARTICLE_ID_COL] if h is not None)
    df = DP.select_rows(df, [h is not None for h in df[DP.DEFAULT_HISTORY_ARTICLE_ID_COL]])
    df = DP.sampling_strategy_wu2019(df, npratio=NP, shuffle=True, with_replacement=True, seed=123)
    df = DP.create_binary_labels_column(df, shuffle=True, seed=123)
    assert all(len(r) == NP + 1 and sum(l) == 1 for r, l in zip(df[INV], df[LAB]))
    art = pq.read_table(REF / "articles.parquet", columns=["article_id"]).to_pydict()["article_id"]
    mapping = {int(a): np.random.default_rng(int(a)).integers(1, 50, 10).tolist() for a in art}
    frame = {k: df[k] for k in (DP.DEFAULT_USER_COL, DP.DEFAULT_HISTORY_ARTICLE_ID_COL, INV, LAB)}
    dl = NRMSDataLoader(behaviors=frame, article_dict=mapping, history_column=DP.DEFAULT_HISTORY_ARTICLE_ID_COL,
                        unknown_representation="zeros", eval_mode=False, batch_size=64)
    (his, pred), y = dl[0]
    assert his.shape == (64, H, 10) and pred.shape == (64, NP + 1, 10) and y.shape == (64, NP + 1) and (y.sum(1) == 1).all()
