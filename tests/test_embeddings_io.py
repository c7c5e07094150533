"""On-disk embedding formats of the reference's extraction scripts.  Files are written with the very
calls the reference uses (pickle.dump of a list of dicts: extract_and_clean_imdb_wiki_faces.py:149-156; np.save of a
list of dicts :232-252; np.save of one dict: extract_faces_from_dataset.py:352-363)."""
import os
import pickle

import numpy as np
import pytest

from oracle import oracle


def _write_fixture(tmp_path, rng, dim=64):
    ref, cand = oracle.make_synthetic(40, 300, dim, seed=3)
    # references: IMDB-WIKI style per-image pickles (some images hold two faces) + an aggregated data.npy
    ref_paths, k = [], 0
    for i in range(12):
        n_faces = 2 if i % 4 == 0 else 1
        faces = [{"det_score": 0.9, "normed_feature": ref[k + j]} for j in range(n_faces)]
        p = str(tmp_path / f"img_{i}.jpg.pkl")
        with open(p, "wb") as f:
            pickle.dump(faces, f)
        ref_paths.append(p)
        k += n_faces
    data = [{"image_path": f"x{j}.jpg", "age": 30, "gender": "M", "feature": ref[j]} for j in range(k, 40)]
    os.makedirs(tmp_path / "agg", exist_ok=True)
    np.save(str(tmp_path / "agg" / "data.npy"), data)
    ref_paths.append(str(tmp_path / "agg" / "data.npy"))
    # candidates: per-media .npy dicts, single vectors and zero-padded concatenations
    cand_paths, k = [], 0
    m = 0
    while k < 300:
        n = int(rng.integers(1, 6))
        n = min(n, 300 - k)
        feat = cand[k:k + n]
        if m % 2 == 0:
            feat = np.concatenate([feat.reshape(-1), np.zeros(3 * dim, np.float32)])      # padded video features
        elif n == 1:
            feat = feat[0]
        p = str(tmp_path / f"media_{m}.npy")
        np.save(p, {"media_id": f"media_{m}", "class_name": "a", "label": 0, "feature": feat.astype(np.float32)})
        cand_paths.append(p)
        k += n
        m += 1
    return ref, cand, ref_paths, cand_paths


def test_loaders_roundtrip(tmp_path):
    from face_detection_and_recognition_b200 import embeddings_io as eio
    rng = np.random.default_rng(0)
    ref, cand, ref_paths, cand_paths = _write_fixture(tmp_path, rng)
    r, rprov = eio.load_embeddings(ref_paths)
    c, cprov = eio.load_embeddings(cand_paths, feature_size=64)
    assert np.array_equal(r, ref) and np.array_equal(c, cand)
    assert len(rprov) == 40 and len(cprov) == 300 and rprov[0] == (ref_paths[0], 0) and rprov[1] == (ref_paths[0], 1)
    x, scores = eio.load_pkl_faces(ref_paths[0], min_det_score=0.95)
    assert x.shape[0] == 0 and scores.shape[0] == 0
    feats, meta = eio.load_data_npy(ref_paths[-1])
    assert feats.shape[1] == 64 and meta[0]["gender"] == "M" and "feature" not in meta[0]
    with pytest.raises(ValueError):
        eio.load_media_npy(cand_paths[0], feature_size=48)
    with pytest.raises(ValueError):
        eio.load_embeddings([str(tmp_path / "x.txt")])


@pytest.mark.gpu
def test_filter_embedding_files(tmp_path, ffr_lib, cuda_dev):
    from face_detection_and_recognition_b200 import embeddings_io as eio
    rng = np.random.default_rng(0)
    ref, cand, ref_paths, cand_paths = _write_fixture(tmp_path, rng)
    keep, idx, val, cprov, rprov = eio.filter_embedding_files(ref_paths, cand_paths, 0.5, feature_size=64)
    ko, io, so = oracle.filter_cosine(ref, cand, 0.5)
    assert np.max(np.abs(val - so)) < 1e-3
    far = np.abs(so - 0.5) > 1e-5
    assert np.array_equal(keep[far], ko[far]) and np.mean(idx == io) > 0.995
    assert len(cprov) == len(keep)
