"""Regenerates the committed fixtures under tests/golden/ (run in the BUILD container only).

 1. ebnerd_sample.json -- a slice of the reference's own parquet fixtures
    (/root/reference/test/data/ebnerd/{behaviors,history}.parquet), reduced the way the reference's
    dataloader test does it (test/dataloader/test_newsrec.py:35-58: history tail(3), binary labels from
    the clicked list, a fabricated 10-token title per article) so the dataloader contract tests can run
    without /root/reference and without polars.
 2. nrms_oracle_case.npz -- seeded inputs/weights and the float64 ORACLE outputs (probabilities, sigmoid
    scores, loss, two gradients).  The reference holds no golden vectors for the model math (parity
    unpinned), so this fixture only pins the oracle against drift; it is labelled as self-generated.
"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
OUT = Path(__file__).resolve().parent


def ebnerd_sample(n_rows=300):
    import pyarrow.parquet as pq

    ref = Path("/root/reference/test/data/ebnerd")
    beh = pq.read_table(ref / "behaviors.parquet", columns=["user_id", "article_ids_inview", "article_ids_clicked"]).to_pydict()
    his = pq.read_table(ref / "history.parquet", columns=["user_id", "article_id_fixed"]).to_pydict()
    art = pq.read_table(ref / "articles.parquet", columns=["article_id"]).to_pydict()
    hist_by_user = {u: h[-3:] for u, h in zip(his["user_id"], his["article_id_fixed"])}
    rows = {"user_id": [], "article_ids_inview": [], "article_ids_clicked": [], "article_id_fixed": [], "labels": []}
    for u, inv, clk in zip(beh["user_id"], beh["article_ids_inview"], beh["article_ids_clicked"]):
        if u not in hist_by_user:
            continue
        rows["user_id"].append(int(u))
        rows["article_ids_inview"].append([int(a) for a in inv])
        rows["article_ids_clicked"].append([int(a) for a in clk])
        rows["article_id_fixed"].append([int(a) for a in hist_by_user[u]])
        rows["labels"].append([1 if a in clk else 0 for a in inv])  # create_binary_labels_column, _behaviors.py:22-107
        if len(rows["user_id"]) == n_rows:
            break
    rng = np.random.default_rng(7)
    used = {a for r in rows["article_ids_inview"] for a in r} | {a for r in rows["article_id_fixed"] for a in r}
    known = [int(a) for a in art["article_id"] if int(a) in used]
    tokens = {str(a): rng.integers(0, 20, 10).tolist() for a in known}
    (OUT / "ebnerd_sample.json").write_text(json.dumps({"behaviors": rows, "article_tokens": tokens}))
    print("ebnerd_sample.json:", len(rows["user_id"]), "impressions,", len(tokens), "articles,",
          len(used) - len(tokens), "ids without an article row (-> unknown index 0)")


def nrms_oracle_case():
    from oracle import nrms_oracle as O

    rng = np.random.default_rng(20240617)
    V, E, nh, dh, att, B, H, C, T = 120, 32, 4, 8, 24, 5, 6, 4, 9
    P = O.init_nrms_params(rng, V, E, nh, dh, att, dtype=np.float64)
    for k in ("news_WQ", "news_WK", "user_WQ", "user_WK"):
        P[k] *= 6.0
    P["news_b"] = rng.standard_normal(att) * 0.05
    his = rng.integers(0, V, (B, H, T)).astype(np.int32)
    pred = rng.integers(0, V, (B, C, T)).astype(np.int32)
    y = np.zeros((B, C), np.float32)
    y[np.arange(B), rng.integers(0, C, B)] = 1
    probs = O.nrms_predict(his, pred, P, nh, dh)
    sig = O.nrms_score(his, pred, P, nh, dh)
    loss, _, G = O.nrms_loss_and_grads(his, pred, y, P, nh, dh, training=True, p_drop=0.2, seed1=11, seed2=22)
    np.savez_compressed(OUT / "nrms_oracle_case.npz", dims=np.array([V, E, nh, dh, att, B, H, C, T]), his=his, pred=pred, y=y,
                        probs=probs, sigmoid=sig, loss=loss, g_news_WV=G["news_WV"], g_user_W=G["user_W"],
                        keep_head=O.dropout_keep_mask(11, 256, 0.2), **{f"P_{k}": v for k, v in P.items()})
    print("nrms_oracle_case.npz: loss", loss)


if __name__ == "__main__":
    ebnerd_sample()
    nrms_oracle_case()
