"""A/B timing of csrc/nn_fast.cu against the kernels it replaces, at the C3 shapes of the U-Net (25 frames, 72x128):
GEGLU of the ds1 / ds2 / ds4 feed-forwards, 3x3 im2col at ds1 (320 and 640 channels), temporal im2col at ds1, temporal
attention at ds1 / ds2.
Prints one JSON line: per kernel, microseconds and achieved GB/s (algorithmic bytes) for both variants.
usage: python tools/bench_nn_fast.py     Results: profiles/r02_first_hw_run.txt."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "guidedvd-3dgs_b200"))
import torch  # noqa: E402

import gvd_native  # noqa: E402
from vc_b200 import ops  # noqa: E402

BF = torch.bfloat16


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    lib = gvd_native.nn()
    F, S = 25, 72 * 128
    cases = {}
    for name, rows, D in (("geglu_ds1", F * S, 1280), ("geglu_ds2", F * S // 4, 2560), ("geglu_ds4", F * S // 16, 5120)):
        h = torch.randn(rows, 2 * D, device="cuda").to(BF)
        out = torch.empty(rows, D, dtype=BF, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        cases[name] = (lambda h=h, out=out, rows=rows, D=D: lib.gvd_geglu(h.data_ptr(), out.data_ptr(), rows, D, st), rows * D * 6)
    for name, C in (("im2col3x3_ds1_c320", 320), ("im2col3x3_ds1_c640", 640)):
        x = torch.randn(F, S, C, device="cuda").to(BF)
        col = torch.empty(F * S, 9 * C, dtype=BF, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        cases[name] = (lambda x=x, col=col, C=C: lib.gvd_im2col3x3_cl(x.data_ptr(), col.data_ptr(), F, 72, 128, C, 1, 0, st), F * S * C * 2 * 10)
    x = torch.randn(F, S, 320, device="cuda").to(BF)
    col = torch.empty(F * S, 3 * 320, dtype=BF, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    cases["im2col_t3_ds1"] = (lambda: lib.gvd_im2col_t3_cl(x.data_ptr(), col.data_ptr(), 1, F, S, 320, st), F * S * 320 * 2 * 4)
    for name, Sx, Hh in (("temporal_attention_ds1", S, 5), ("temporal_attention_ds2", S // 4, 10)):
        qkv = [torch.randn(F, Sx, Hh * 64, device="cuda").to(BF) for _ in range(3)]
        o = torch.empty_like(qkv[0])
        st = torch.cuda.current_stream().cuda_stream
        cases[name] = (lambda qkv=qkv, o=o, Sx=Sx, Hh=Hh: lib.gvd_temporal_attention(qkv[0].data_ptr(), qkv[1].data_ptr(), qkv[2].data_ptr(),
                                                                                     o.data_ptr(), 1, F, Sx, Hh, 0.125, st), F * Sx * Hh * 64 * 2 * 4)
    res = {}
    for name, (fn, nbytes) in cases.items():
        row = {}
        for fast in (0, 2):
            lib.gvd_nn_set_fast(fast)
            us = timed(fn)
            row["fast" if fast else "base"] = {"us": round(us, 1), "GBps": round(nbytes / us / 1e3, 1)}
        res[name] = row
    lib.gvd_nn_set_fast(1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
