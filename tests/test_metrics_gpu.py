"""GPU: on-device AUROC / AP (aaclip_minmax, aaclip_score_histogram, aaclip_auc_ap) against scikit-learn -- the
implementation the reference's metrics_eval calls (forward_utils.py:288-296) -- and metrics_eval_device against the
host mirror of metrics_eval on the same arrays. north_star: AUROC identical to 3 decimals."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import aaclip_b200  # noqa: F401
    from aaclip_b200 import metrics
    return metrics


@pytest.mark.parametrize("n,sep,frac", [(2_000_000, 1.0, 0.02), (300_000, 0.1, 0.5), (5_000, 3.0, 0.3)])
def test_auc_ap_match_sklearn(M, n, sep, frac):
    from sklearn.metrics import average_precision_score, roc_auc_score
    rng = np.random.default_rng(n)
    y = (rng.random(n) < frac).astype(np.uint8)
    s = (rng.normal(size=n) + sep * y).astype(np.float32)
    s[: n // 10] = np.round(s[: n // 10], 1)                  # plenty of exact ties
    sn = (s - s.min()) / (s.max() - s.min())
    auc, ap = M.roc_auc_ap(torch.from_numpy(s).cuda(), torch.from_numpy(y).cuda(), normalize=True)
    assert abs(auc - roc_auc_score(y, sn)) < 2e-6
    assert abs(ap - average_precision_score(y, sn)) < 2e-6


def test_degenerate_inputs(M):
    s = torch.rand(1000, device="cuda")
    auc, ap = M.roc_auc_ap(s, torch.zeros(1000, device="cuda"))           # no positives
    assert auc == 0 and ap == 0
    auc, ap = M.roc_auc_ap(torch.full((1000,), 0.5, device="cuda"), (torch.arange(1000, device="cuda") % 2))
    assert abs(auc - 0.5) < 1e-12 and abs(ap - 0.5) < 1e-12               # all tied
    mm = M.minmax([torch.tensor([3.0, -2.0], device="cuda"), torch.tensor([7.5], device="cuda")])
    assert mm.cpu().tolist() == [-2.0, 7.5]                               # streaming merge over chunks
    with pytest.raises(ValueError):
        M.histogram([s], [torch.zeros(999, device="cuda")], None)


@pytest.mark.parametrize("domain", ["Industrial", "Medical"])
def test_metrics_eval_device_matches_host_mirror(M, domain):
    from aaclip_b200 import forward_utils as fu
    rng = np.random.default_rng(7)
    n, S = 12, 96
    masks = (rng.random((n, 1, S, S)) > 0.93).astype(np.float32)
    labels = (masks.reshape(n, -1).max(1) > 0).astype(np.int64)
    labels[:3] = 0
    preds = (rng.normal(size=(n, S, S)) + 1.5 * masks[:, 0]).astype(np.float32) * 3.0 - 1.0
    scores = rng.random(n).astype(np.float32)
    want = fu.metrics_eval(masks, labels, preds, scores, "carpet", domain)
    # device version, maps delivered in ragged batches as an evaluation loop would
    cuts = [0, 5, 9, 12]
    got = M.metrics_eval_device([torch.from_numpy(masks[a:b]).cuda() for a, b in zip(cuts, cuts[1:])],
                                torch.from_numpy(labels).cuda(),
                                [torch.from_numpy(preds[a:b]).cuda() for a, b in zip(cuts, cuts[1:])],
                                torch.from_numpy(scores).cuda(), "carpet", domain)
    for k in ("pixel AUC", "pixel AP", "image AUC", "image AP"):
        assert abs(got[k] - want[k]) <= 0.0101, (k, got[k], want[k])       # 4-decimal rounding of values equal to ~1e-6
    assert got["class name"] == "carpet"
