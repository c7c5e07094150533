"""Throughput of the device-resident face tracker (ops.FaceGallery = ffr_first_match_stream; SURVEY §8 row f4: the reference's
Net.check_if_face_exists + add_face, extract_and_label_faces_from_dataset.py:101-121).

A stream of noisy re-appearances of ``n_id`` identities is matched in ONE launch (queries depend on each other through the
gallery, so the kernel walks them in order: the figure is the latency of one query).  CUDA events around the launch, inputs
resident; the CPU line is the reference's scan restated in oracle.first_match_scan (pure NumPy per pair, one thread) on a
bounded prefix.  Prints one JSON line per configuration.

    python tools/bench_tracker.py [--queries 20000]
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_and_recognition_b200 import ops  # noqa: E402


def stream(n_id, n_q, dim, seed=0):
    rng = np.random.default_rng(seed)
    ids = rng.standard_normal((n_id, dim)).astype(np.float32)
    ids /= np.linalg.norm(ids, axis=1, keepdims=True)
    who = rng.integers(0, n_id, n_q)
    q = ids[who] + np.float32(0.35 / np.sqrt(dim)) * rng.standard_normal((n_q, dim)).astype(np.float32)
    box = rng.uniform(0, 200, (n_q, 4)).astype(np.float32)
    box[:, 2:] += box[:, :2] + 20
    return q, box


def main():
    n_q = int(sys.argv[sys.argv.index("--queries") + 1]) if "--queries" in sys.argv else 20000
    cpu = "--no-cpu" not in sys.argv
    dev = torch.device("cuda:0")
    # an idle B200 sits at 120 MHz and takes ~0.6 s of load to reach its boost clock: a 40 ms single-CTA kernel timed cold would
    # be timed at whatever clock the ramp has reached -- keep the GPU busy first, and between configurations
    def warm(seconds=1.0):
        a = torch.randn(4096, 4096, device=dev)
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for _ in range(20):
                a @ a
            torch.cuda.synchronize()
    warm(1.5)
    for dim, n_id, cap in [(128, 16, 64), (128, 128, 256), (128, 300, 384), (512, 64, 96), (256, 128, 160), (128, 128, 4096),
                           (128, 1024, 4096), (512, 1024, 4096)]:
        q, box = stream(n_id, n_q, dim, seed=dim + n_id)
        qd, bd = torch.from_numpy(q).to(dev), torch.from_numpy(box).to(dev)
        times = []
        warm(0.3)
        for rep in range(4):
            gal = ops.FaceGallery(dim, capacity=cap, metric="cosine")
            gal.match(qd[:64], bd[:64])                      # first launch: module load, attribute
            gal.clear()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            found, _ = gal.match(qd, bd)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = float(np.median(times[1:]))
        line = {"dim": dim, "identities": n_id, "capacity": cap, "queries": n_q, "gallery_at_end": len(gal),
                "resident_in_smem": (8 + cap) * (dim + 4) * 4 <= 200 * 1024, "ms": round(ms, 3),
                "us_per_query": round(1e3 * ms / n_q, 3), "queries_per_s": round(n_q / ms * 1e3),
                "found_fraction": round(float(found.mean()), 4)}
        if cpu:
            sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
            from oracle import oracle
            n_cpu = min(n_q, 400)
            gallery = []
            t0 = time.perf_counter()
            for x in q[:n_cpu]:
                f, pos = oracle.first_match_scan(gallery, x, oracle.METRIC_COSINE)
                if f:
                    gallery[pos] = x
                else:
                    gallery.append(x)
            dt = time.perf_counter() - t0
            line["cpu_port_queries_per_s"] = round(n_cpu / dt)
            line["cpu_sample"] = f"first {n_cpu} queries, oracle.first_match_scan, 1 thread"
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
