"""A/B a variant build of the library (python -m whvi_b200.build --variant NAME -D...) against the
product library: same inputs through whvi_layer_fwd_f32 / whvi_layer_bwd_f32 of both, outputs
compared, both timed.      python tools/lab_ab.py tools/lab/libwhvi_b200_NAME.so [D ...]"""
import ctypes
import sys
from ctypes import c_int64, c_size_t, c_void_p
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def load(path):
    lib = ctypes.CDLL(str(path))
    lib.whvi_layer_bwd_workspace_bytes.argtypes = [c_int64, c_int64, c_int64, ctypes.POINTER(c_size_t)]
    lib.whvi_layer_bwd_f32.argtypes = [c_void_p, c_int64] + [c_void_p] * 10 + [c_size_t, c_int64, c_int64, c_int64, c_void_p]
    lib.whvi_layer_fwd_f32.argtypes = [c_void_p, c_int64] + [c_void_p] * 5 + [c_int64, c_int64, c_int64, c_void_p]
    lib.whvi_last_error.restype = ctypes.c_char_p
    return lib


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    b.synchronize()
    return a.elapsed_time(b) / iters


def main():
    libs = {"product": load(ROOT / "whvi_b200" / "libwhvi_b200.so"), "variant": load(sys.argv[1])}
    dev = torch.device("cuda:0")
    for D in [int(a) for a in sys.argv[2:]] or [4096]:
        S, B = 16, (1 << 27) // (16 * D)
        x, dy = torch.randn(S, B, D, device=dev), torch.randn(S, B, D, device=dev)
        g, s1, s2 = torch.randn(S, D, device=dev), torch.randn(D, device=dev), torch.randn(D, device=dev)
        res = {}
        for name, lib in libs.items():
            y, dx = torch.empty_like(x), torch.empty_like(x)
            dg, ds1, ds2 = torch.empty_like(g), torch.empty_like(s1), torch.empty_like(s2)
            need = c_size_t(0)
            assert lib.whvi_layer_bwd_workspace_bytes(S, B, D, ctypes.byref(need)) == 0
            ws = torch.empty(need.value, dtype=torch.uint8, device=dev)

            def fwd():
                rc = lib.whvi_layer_fwd_f32(x.data_ptr(), B * D, g.data_ptr(), s1.data_ptr(), s2.data_ptr(), None, y.data_ptr(),
                                            S, B, D, None)
                assert rc == 0, lib.whvi_last_error()

            def bwd():
                rc = lib.whvi_layer_bwd_f32(x.data_ptr(), B * D, dy.data_ptr(), g.data_ptr(), s1.data_ptr(), s2.data_ptr(),
                                            dx.data_ptr(), dg.data_ptr(), ds1.data_ptr(), ds2.data_ptr(), None, ws.data_ptr(),
                                            ws.numel(), S, B, D, None)
                assert rc == 0, lib.whvi_last_error()

            res[name] = (timed(fwd), timed(bwd), y, dx, dg, ds1, ds2)
        p, v = res["product"], res["variant"]
        worst = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(v[2:], p[2:]))
        n = S * B * D
        print(f"D={D}: fwd {p[0]:.3f} -> {v[0]:.3f} ms ({8.0 * n / v[0] / 1e6:.0f} GB/s)   bwd {p[1]:.3f} -> {v[1]:.3f} ms "
              f"({12.0 * n / v[1] / 1e6:.0f} GB/s)   max rel diff of outputs {worst:.2e}")


if __name__ == "__main__":
    main()
