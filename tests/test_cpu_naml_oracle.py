"""CPU tests of the NAML oracle: Conv1D 'same' against torch's conv1d, analytic backward against autograd."""
import numpy as np
import pytest
import torch
import torch.nn.functional as Fn

from oracle import naml_oracle as NA, nrms_oracle as O


def small(rng=None, V=40, E=8, F=12, att=6, window=3, B=3, H=4, C=3, T=5, Tb=7, ncat=9, dcat=3):
    rng = rng or np.random.default_rng(0)
    P = NA.init_naml_params(rng, V, E, F, att, window, ncat, dcat, ncat + 2, dcat + 1, dtype=np.float64)
    for k in P:
        if k.endswith("_b") or k.endswith("convb") or k.endswith("denseb"):
            P[k] = rng.standard_normal(P[k].shape) * 0.1
    arrays = [rng.integers(0, V, (B, H, T)), rng.integers(0, V, (B, H, Tb)), rng.integers(0, ncat, (B, H, 1)),
              rng.integers(0, ncat + 2, (B, H, 1)), rng.integers(0, V, (B, C, T)), rng.integers(0, V, (B, C, Tb)),
              rng.integers(0, ncat, (B, C, 1)), rng.integers(0, ncat + 2, (B, C, 1))]
    y = np.zeros((B, C))
    y[np.arange(B), rng.integers(0, C, B)] = 1
    art, B, H, C = NA.pack_inputs(*arrays)
    return P, art, y, (B, H, C, T, Tb)


@pytest.mark.parametrize("window", [1, 2, 3, 4, 5])
def test_conv1d_same_matches_torch(window):
    """Keras Conv1D(padding='same') == cross-correlation with (w-1)//2 zeros left, the rest right (naml.py:159-166)."""
    rng = np.random.default_rng(window)
    X, Wc, bc = rng.standard_normal((3, 7, 4)), rng.standard_normal((window, 4, 5)), rng.standard_normal(5)
    y, _ = NA.conv1d_same_fwd(X, Wc, bc)
    xt = torch.tensor(X).permute(0, 2, 1)  # [N, E, L]
    padl = (window - 1) // 2
    xt = Fn.pad(xt, (padl, window - 1 - padl))
    ref = Fn.conv1d(xt, torch.tensor(Wc).permute(2, 1, 0), torch.tensor(bc)).permute(0, 2, 1).relu().numpy()
    np.testing.assert_allclose(y, ref, rtol=1e-12, atol=1e-13)


def torch_naml_loss(art, y, Pt, B, H, C, T, Tb, keeps, p):
    def att(X, W, b, q):
        a = torch.tanh(X @ W + b) @ q
        e = torch.exp(a[..., 0])
        w = e / (e.sum(-1, keepdim=True) + 1e-7)
        return (w[..., None] * X).sum(-2)

    def text(tok, v, k1, k2):
        X = Pt["table"][torch.tensor(tok)]
        s = 1.0 / (1.0 - p)
        X = X * k1 * s
        Wc = Pt[f"{v}_convW"]
        w = Wc.shape[0]
        padl = (w - 1) // 2
        xt = Fn.pad(X.permute(0, 2, 1), (padl, w - 1 - padl))
        yv = Fn.conv1d(xt, Wc.permute(2, 1, 0), Pt[f"{v}_convb"]).permute(0, 2, 1).relu()
        return att(yv * k2 * s, Pt[f"{v}_W"], Pt[f"{v}_b"], Pt[f"{v}_q"])

    def cat(ids, v):
        return (Pt[f"{v}_emb"][torch.tensor(ids)] @ Pt[f"{v}_denseW"] + Pt[f"{v}_denseb"]).relu()

    views = torch.stack([text(art[:, :T], "title", keeps[0], keeps[1]), text(art[:, T:T + Tb], "body", keeps[2], keeps[3]),
                         cat(art[:, T + Tb], "vert"), cat(art[:, T + Tb + 1], "subvert")], dim=1)
    n_all = att(views, Pt["news_W"], Pt["news_b"], Pt["news_q"])
    Fd = n_all.shape[-1]
    u = att(n_all[:B * H].reshape(B, H, Fd), Pt["user_W"], Pt["user_b"], Pt["user_q"])
    z = torch.einsum("bcd,bd->bc", n_all[B * H:].reshape(B, C, Fd), u)
    return -(torch.tensor(y) * torch.log_softmax(z, -1)).sum(-1).mean()


@pytest.mark.parametrize("window", [3, 4])
def test_naml_analytic_backward_matches_autograd(window):
    P, art, y, (B, H, C, T, Tb) = small(window=window)
    p, seeds = 0.2, (5, 6, 7, 8)
    loss, prob, G = NA.naml_loss_and_grads(art, B, H, C, y, P, T, Tb, p_drop=p, seeds=seeds)
    N, E, F = art.shape[0], P["table"].shape[1], P["title_convb"].shape[0]
    keeps = [torch.tensor(O.dropout_keep_mask(s, N * L * W, p).reshape(N, L, W), dtype=torch.float64)
             for s, L, W in ((5, T, E), (6, T, F), (7, Tb, E), (8, Tb, F))]
    Pt = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in P.items()}
    l = torch_naml_loss(art, y, Pt, B, H, C, T, Tb, keeps, p)
    l.backward()
    assert abs(float(l.detach()) - float(loss)) < 1e-10
    for k in P:
        np.testing.assert_allclose(G[k], Pt[k].grad.numpy(), rtol=1e-8, atol=1e-13, err_msg=k)
    np.testing.assert_allclose(prob.sum(-1), 1.0, rtol=1e-12)


def test_naml_param_shapes_and_count():
    """naml_dummy.py config: table 1000x100, filter 400, window 3, att 200, 100x10 category tables."""
    P = NA.init_naml_params(np.random.default_rng(0), 1000, 100, 400, 200, 3, 100, 10, 100, 10)
    assert P["title_convW"].shape == (3, 100, 400) and P["vert_emb"].shape == (100, 10)
    n = sum(v.size for v in P.values())
    want = 1000 * 100 + 2 * (3 * 100 * 400 + 400 + 400 * 200 + 200 + 200) + 2 * (100 * 10 + 10 * 400 + 400) + 2 * (400 * 200 + 200 + 200)
    assert n == want
    assert list(P) == NA.NAML_PARAM_ORDER


def test_out_of_range_ids_read_zero_rows():
    P, art, y, (B, H, C, T, Tb) = small()
    art2 = art.copy()
    art2[0, 0] = 10_000          # title token outside the table
    art2[1, T + Tb] = 10_000     # vert id outside its table
    z, _ = NA.naml_forward(art2, B, H, C, P, T, Tb)
    P2 = {k: v.copy() for k, v in P.items()}
    P2["table"] = np.concatenate([P["table"], np.zeros((10_001 - P["table"].shape[0], P["table"].shape[1]))])
    P2["vert_emb"] = np.concatenate([P["vert_emb"], np.zeros((10_001 - P["vert_emb"].shape[0], P["vert_emb"].shape[1]))])
    z2, _ = NA.naml_forward(art2, B, H, C, P2, T, Tb)
    np.testing.assert_allclose(z, z2, rtol=1e-12)
