#!/usr/bin/env python
"""GPU diagnostic: raw tcgen05 score tiles against a torch fp32 matmul of the same fp16 rows, then quick timings.
Run on the GPU box:  python tools/diag_mma.py [--perf]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from face_detection_and_recognition_b200 import _lib, ops  # noqa: E402


def raw(n_ref, n_cand, dim):
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    ref = torch.randn(n_ref, dim, device="cuda", generator=g)
    cand = torch.randn(n_cand, dim, device="cuda", generator=g)
    r16 = ops.l2norm_rows(ref, want_f16=True, want_f32=False)["f16"]
    c16 = ops.l2norm_rows(cand, want_f16=True, want_f32=False)["f16"]
    ld = r16.shape[1]
    scores = torch.full((n_cand, n_ref), float("nan"), device="cuda")
    keep = torch.empty(n_cand, dtype=torch.uint8, device="cuda")
    idx = torch.empty(n_cand, dtype=torch.int32, device="cuda")
    val = torch.empty(n_cand, dtype=torch.float32, device="cuda")
    ws = torch.zeros(4096 + 64 * n_cand, dtype=torch.uint8, device="cuda")
    rc = lib.ffr_debug_mma_scores(r16.data_ptr(), n_ref, c16.data_ptr(), n_cand, ld, 0.5, keep.data_ptr(),
                                  idx.data_ptr(), val.data_ptr(), scores.data_ptr(), ws.data_ptr(), ws.numel(),
                                  torch.cuda.current_stream().cuda_stream)
    if rc != 0:
        print("  rc", rc, lib.ffr_last_error().decode())
        return False
    torch.cuda.synchronize()
    want = c16.float() @ r16.float().T
    err = (scores - want).abs()
    nan = torch.isnan(scores).sum().item()
    e = torch.nan_to_num(err, nan=9.0)
    print(f"  [{n_ref}x{n_cand}x{dim}] nan={nan} max_err={e.max().item():.3e} mean_err={e.mean().item():.3e} "
          f"val_err={(val - want.max(1).values).abs().max().item():.3e} "
          f"idx_match={(idx.long() == want.argmax(1)).float().mean().item():.4f}")
    if e.max().item() > 1e-4:
        bad = (e > 1e-4)
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()
        print("   bad rows (first 8):", rows, " bad cols: count", cols.numel(), "first", cols[:16].tolist())
        print("   got[0,:8]", scores[0, :8].tolist())
        print("   want[0,:8]", want[0, :8].tolist())
        # does the result look like a permutation of K chunks / rows?
        for shift in (8, 16, 32, 64):
            if n_ref > shift:
                print(f"   err vs want shifted by {shift} cols:", (scores[:, :-shift] - want[:, shift:]).abs().max().item())
        return False
    return True


def perf(n_ref, n_cand, dim, flags=0, iters=5):
    g = torch.Generator(device="cuda").manual_seed(2)
    ref = torch.nn.functional.normalize(torch.randn(n_ref, dim, device="cuda", generator=g))
    cand = torch.nn.functional.normalize(torch.randn(n_cand, dim, device="cuda", generator=g))
    cand[::2] = torch.nn.functional.normalize(ref[torch.randint(0, n_ref, (cand[::2].shape[0],), device="cuda")] * 0.8
                                              + cand[::2] * 0.6)
    res = ops.face_filter(ref, cand, 0.5, flags=flags, want_stats=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        ops.face_filter(ref, cand, 0.5, flags=flags, out=(res.keep, res.best_idx, res.best_val))
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = min(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    pairs = n_ref * n_cand
    print(f"  perf [{n_ref}x{n_cand}x{dim}] flags={flags}: {ms:.3f} ms  {pairs / ms / 1e6:.1f} Gpairs/s  "
          f"{2 * pairs * dim / ms / 1e9:.1f} TFLOP/s  stats={res.stats} keep={res.keep.float().mean().item():.3f}")


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), _lib.load().ffr_build_info().decode())
    ok = True
    for shape in [(256, 128, 64), (256, 128, 128), (512, 256, 512), (100, 77, 128), (1000, 333, 192), (4096, 4096, 256)]:
        ok = raw(*shape) and ok
    print("RAW", "OK" if ok else "FAILED")
    if "--perf" in sys.argv and ok:
        perf(1000, 100_000, 128)
        perf(1000, 100_000, 128, flags=ops.FLAG_NO_RECHECK)
        perf(10_000, 1_000_000, 512)
        perf(10_000, 1_000_000, 512, flags=ops.FLAG_NO_RECHECK)
        perf(100_000, 1_250_000, 128, flags=ops.FLAG_NO_RECHECK, iters=2)
        perf(1, 10_000_000, 128)
