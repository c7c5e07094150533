"""Full-size runs (BASELINE.json configs[2] and the reference's literal mode at 10M rows) checked through
size-independent properties plus an oracle sample: planted matches are recovered, results are equivariant under
candidate permutation, invariant under positive row scaling, and consistent under a split of the reference axis."""
import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(ffr_lib, cuda_dev):
    from face_detection_and_recognition_b200 import ops as _ops
    return _ops


def _make(n_ref, n_cand, dim, seed=42):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ref = torch.nn.functional.normalize(torch.randn(n_ref, dim, device="cuda", generator=g))
    noise = torch.nn.functional.normalize(torch.randn(n_cand, dim, device="cuda", generator=g))
    k = torch.randint(0, n_ref, (n_cand,), device="cuda", generator=g)
    c = torch.rand(n_cand, device="cuda", generator=g) * 0.4 + 0.55
    planted = torch.nn.functional.normalize(c[:, None] * ref[k] + torch.sqrt(1 - c * c)[:, None] * noise)
    even = (torch.arange(n_cand, device="cuda") % 2 == 0)
    cand = torch.where(even[:, None], planted, noise)
    return ref, cand, k, even


def test_config2_fullsize(ops):
    n_ref, n_cand, dim, thr = 10_000, 1_000_000, 512, 0.5
    ref, cand, k, even = _make(n_ref, n_cand, dim)
    res = ops.face_filter(ref, cand, thr, want_stats=True)
    torch.cuda.synchronize()
    assert res.stats["path"] == "tcgen05"
    keep, idx, val = res.keep, res.best_idx, res.best_val

    # (1) planted rows recover their reference; the reported similarity is the true fp32 cosine within 1e-3
    true_cos = (cand[even] * ref[k[even]]).sum(1)
    assert torch.equal(idx[even].long(), k[even])
    assert (val[even] - true_cos).abs().max().item() < 1e-3
    clear = (true_cos - thr).abs() > 1e-3                 # the nominal cos >= 0.55 can dip under 0.5: noise is not orthogonal
    assert torch.equal(keep[even][clear].bool(), (true_cos >= thr)[clear])
    assert keep[even].float().mean().item() > 0.95
    # noise rows: best over 10k random references stays far below the threshold
    assert not keep[~even].any() and val[~even].max().item() < 0.35
    # (2) keep is the threshold test of best_val everywhere
    assert torch.equal(keep.bool(), val >= thr)

    # (3) oracle on a random sample of rows (all references)
    sel = torch.randperm(n_cand, device="cuda")[:4096]
    ko, io, so = oracle.filter_cosine(ref.cpu().numpy(), cand[sel].cpu().numpy(), thr)
    _, _, s64 = oracle.filter_cosine(ref.cpu().numpy(), cand[sel].cpu().numpy(), thr, dtype=np.float64)
    assert np.max(np.abs(val[sel].cpu().numpy() - so)) < 1e-3
    assert np.mean(idx[sel].cpu().numpy() == io) > 0.9995
    far = np.abs(s64 - thr) > 1e-6
    assert np.array_equal(keep[sel].cpu().numpy()[far], ko[far])

    # (4) candidate-permutation equivariance (bit exact: a row's result does not depend on where it sits)
    m = 200_000
    perm = torch.randperm(m, device="cuda")
    a = ops.face_filter(ref, cand[:m], thr)
    b = ops.face_filter(ref, cand[:m][perm].contiguous(), thr)
    assert torch.equal(b.best_idx, a.best_idx[perm]) and torch.equal(b.keep, a.keep[perm])
    assert torch.equal(b.best_val, a.best_val[perm])

    # (5) positive row scaling changes nothing (cosine): same indices and masks
    scale = torch.rand(m, 1, device="cuda") * 9 + 0.5
    c = ops.face_filter(ref, (cand[:m] * scale).contiguous(), thr)
    assert (c.best_val - a.best_val).abs().max().item() < 2e-4
    same = c.best_idx == a.best_idx
    assert same.float().mean().item() > 0.9999
    stable = (a.best_val - thr).abs() > 1e-3
    assert torch.equal(c.keep[stable], a.keep[stable])

    # (6) splitting the reference axis and merging (max, first index on ties) reproduces the unsplit result
    lo = ops.face_filter(ref[:5000].contiguous(), cand[:m], thr)
    hi = ops.face_filter(ref[5000:].contiguous(), cand[:m], thr, ref_index_base=5000)
    take_hi = hi.best_val > lo.best_val
    mval = torch.where(take_hi, hi.best_val, lo.best_val)
    midx = torch.where(take_hi, hi.best_idx, lo.best_idx)
    assert (mval - a.best_val).abs().max().item() < 2e-4
    clear = (hi.best_val - lo.best_val).abs() > 1e-3
    assert torch.equal(midx[clear], a.best_idx[clear])


def test_literal_mode_fullsize(ops):
    """The reference's literal mode (one mean vector, Euclid keep test, filter_faces_using_reference.py:186-189) on 10M
    rows: against a plain torch fp32 evaluation of the same expression and the oracle on a sample."""
    n_cand, dim, thr = 10_000_000, 128, 1.2
    ref, cand, k, even = _make(1, n_cand, dim, seed=7)
    res = ops.face_filter(ref, cand, thr, metric="euclid", want_stats=True)
    assert res.stats["path"] == "fp32"
    d = torch.linalg.vector_norm(cand - ref, dim=1)
    assert (res.best_val - d).abs().max().item() < 2e-6
    clear = (d - thr).abs() > 1e-5
    assert torch.equal(res.keep.bool()[clear], (d <= thr)[clear])
    assert (res.best_idx == 0).all()
    sel = torch.randperm(n_cand, device="cuda")[:100_000]
    ko = oracle.euclid_keep_literal(cand[sel].cpu().numpy()[:5000], ref.cpu().numpy(), np.float32(thr))
    near = (d[sel][:5000] - thr).abs().cpu().numpy() < 1e-5
    assert np.array_equal(res.keep[sel][:5000].cpu().numpy()[~near], ko[~near])
    assert 0.45 < res.keep.float().mean().item() < 0.55
