"""GPU tests of the "next" rows: batched reference statistics (K5 over all classes) and the streaming
first-match gallery scan, both against outputs of the reference itself (tests/golden)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(ffr_lib, cuda_dev):
    from face_detection_and_recognition_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("name", ["ref_main_facenetlike.npz", "ref_main_mobilefacenet.npz"])
def test_batched_ref_stats_match_reference_main(ops, golden_dir, name):
    """One launch for all classes == the (mean, thres) the reference's main() computed class by class (:85-99)."""
    g = np.load(os.path.join(golden_dir, name))
    feats = [g[f"c{c}_ref_feat"].reshape(-1, g[f"c{c}_ref_feat"].shape[-1]) for c in range(3)]
    counts = [f.shape[0] for f in feats]
    mean, thres = ops.ref_mean_and_thres_batched(torch.from_numpy(np.concatenate(feats)).cuda(), counts)
    for c in range(3):
        np.testing.assert_allclose(mean[c:c + 1].cpu().numpy(), g[f"c{c}_mu"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(float(thres[c]), float(g[f"c{c}_thres"]), rtol=1e-6)
        m1, t1 = ops.ref_mean_and_thres(torch.from_numpy(feats[c]).cuda())
        assert torch.equal(m1, mean[c:c + 1]) and float(t1) == float(thres[c])


def test_batched_ref_stats_ragged(ops):
    rng = np.random.default_rng(3)
    counts = [1, 32, 7, 100]
    x = rng.standard_normal((sum(counts), 200)).astype(np.float32) * 5
    mean, thres = ops.ref_mean_and_thres_batched(torch.from_numpy(x).cuda(), counts)
    o = 0
    for c, n in enumerate(counts):
        mu, th = oracle.ref_mean_vec_and_thres(x[o:o + n][:, None, :], n)
        np.testing.assert_allclose(mean[c].cpu().numpy(), mu[0], rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(float(thres[c]), float(th), rtol=2e-6, atol=1e-6)
        o += n


@pytest.mark.parametrize("tag,metric,dim", [("mfn", "euclid", 512), ("reid", "cosine", 256)])
def test_first_match_stream_matches_reference_tracker(ops, golden_dir, tag, metric, dim):
    """FaceGallery.match == the (found, faceid) sequence Net.check_if_face_exists / add_face produced in the
    reference (extract_and_label_faces_from_dataset.py:101-121), for the Euclid (MobileFaceNet) and cosine branches."""
    g = np.load(os.path.join(golden_dir, "label_scan_ref.npz"))
    feats, bboxes = g[f"{tag}_feats"], g[f"{tag}_bboxes"]
    want_found, want_id = g[f"{tag}_found"].astype(bool), g[f"{tag}_faceid"]
    gal = ops.FaceGallery(dim, capacity=256, metric=metric, normal_thres=1.0, harsh_thres=0.72)
    # all queries in one launch (sequential inside the kernel) ...
    found, faceid = gal.match(feats, bboxes)
    assert np.array_equal(found, want_found), np.flatnonzero(found != want_found)
    assert np.array_equal(faceid[want_found], want_id[want_found])
    assert len(gal) == int((~want_found).sum())
    # ... and one query per call give the same trajectory
    gal2 = ops.FaceGallery(dim, capacity=256, metric=metric)
    f2 = np.array([gal2.match(feats[i], bboxes[i])[0][0] for i in range(len(feats))])
    assert np.array_equal(f2, want_found)


def _tracker_stream(n_id, n_q, dim, seed):
    """Noisy re-appearances of n_id identities (cosine ~0.94 to their identity, ~0 to the others), random order."""
    rng = np.random.default_rng(seed)
    ids = rng.standard_normal((n_id, dim)).astype(np.float32)
    ids /= np.linalg.norm(ids, axis=1, keepdims=True)
    who = rng.integers(0, n_id, n_q)
    q = ids[who] + np.float32(0.35 / np.sqrt(dim)) * rng.standard_normal((n_q, dim)).astype(np.float32)
    return (q * rng.uniform(0.5, 3.0, (n_q, 1))).astype(np.float32), who


@pytest.mark.parametrize("metric,dim,n_id,n_q", [("cosine", 128, 40, 1500), ("cosine", 512, 70, 600), ("euclid", 256, 150, 900),
                                                 ("cosine", 100, 300, 1200)])
def test_first_match_stream_resident_and_l2_galleries_agree(ops, metric, dim, n_id, n_q):
    """Round 2: a gallery whose whole capacity fits in shared memory lives there for the launch (write-through), larger ones are
    scanned out of L2; 128 entries per round, the next query prefetched.  Both forms, one launch or many, give the trajectory
    of the reference's scan restated in oracle.first_match_scan (+ its in-place update / append, :113-121)."""
    q, _ = _tracker_stream(n_id, n_q, dim, seed=dim + n_id)
    if metric == "euclid":
        q /= np.linalg.norm(q, axis=1, keepdims=True)            # MobileFaceNet embeddings are unit vectors: dist < 0.72 <=> same face
    small_cap = n_id + 8
    assert (8 + small_cap) * (dim + 4) * 4 <= 200 * 1024   # the resident form
    res = []
    for cap, chunks in ((small_cap, 1), (4096, 1), (small_cap, 7)):
        gal = ops.FaceGallery(dim, capacity=cap, metric=metric)
        out = [gal.match(part) for part in np.array_split(q, chunks)]
        res.append((np.concatenate([o[0] for o in out]), np.concatenate([o[1] for o in out]), len(gal),
                    gal.feat[:len(gal)].cpu().numpy()))
    for r in res[1:]:
        assert np.array_equal(r[0], res[0][0]) and np.array_equal(r[1], res[0][1]) and r[2] == res[0][2]
        assert np.array_equal(r[3], res[0][3])                   # the global copy of the gallery is current after every launch
    gallery, want_found, want_id = [], [], []
    m = oracle.METRIC_EUCLID if metric == "euclid" else oracle.METRIC_COSINE
    for x in q[:400]:
        found, pos = oracle.first_match_scan(gallery, x, m)
        if found:
            gallery[pos] = x
        else:
            gallery.append(x)
            pos = len(gallery) - 1
        want_found.append(found)
        want_id.append(pos + 1)
    assert np.array_equal(res[0][0][:400], np.array(want_found)) and np.array_equal(res[0][1][:400], np.array(want_id))
    assert res[0][2] <= n_id + n_id // 4 + 8                     # (nearly) every re-appearance was recognised


def test_first_match_capacity_and_no_bbox(ops):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((10, 64)).astype(np.float32)
    gal = ops.FaceGallery(64, capacity=4, metric="cosine")
    found, faceid = gal.match(x[:4])                 # four unrelated faces fill the gallery
    assert not found.any() and list(faceid) == [1, 2, 3, 4]
    found, faceid = gal.match(x[2] * 3.0)            # a scaled copy: cosine distance 0 < harsh threshold
    assert found[0] and faceid[0] == 3
    with pytest.raises(RuntimeError):
        gal.match(x[5])


@pytest.mark.parametrize("dim,n", [(128, 32), (256, 17), (512, 32), (100, 9), (1024, 5), (130, 12), (384, 40)])
def test_farthest_reference_passes_its_own_threshold(ops, dim, n):
    """The reference takes the threshold (:88-99) and the keep test (:189) from the same np.linalg.norm, so the reference
    image that DEFINES the threshold always passes `<= thres` when it is also a candidate.  K5 computes its distances with
    the arithmetic of the streaming filter (same per-lane element order, same reduction tree): bit-identical, no 1-ulp flip."""
    rng = np.random.default_rng(dim + n)
    for trial in range(20):
        x = torch.from_numpy((rng.standard_normal((n, dim)) * rng.uniform(0.5, 12)).astype(np.float32)).cuda()
        for mean, thres in (ops.ref_mean_and_thres(x), tuple(t[0:1] if t.dim() == 2 else t[0] for t in ops.ref_mean_and_thres_batched(x, [n]))):
            res = ops.face_filter(mean.reshape(1, -1), x, float(thres), metric="euclid")
            assert bool(res.keep.all()), (trial, float(thres), float(res.best_val.max()))
            assert float(res.best_val.max()) == float(thres)
