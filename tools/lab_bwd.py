"""Time whvi_layer_bwd_f32 from an arbitrary build of the library (A/B experiments)."""
import ctypes, sys
from ctypes import c_void_p, c_int64, c_size_t, c_int
import torch
lib = ctypes.CDLL(sys.argv[1])
D = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
S, B = 16, (1 << 28) // (16 * D)
lib.whvi_layer_bwd_workspace_bytes.argtypes = [c_int64, c_int64, c_int64, ctypes.POINTER(c_size_t)]
lib.whvi_layer_bwd_f32.argtypes = [c_void_p, c_int64] + [c_void_p] * 10 + [c_size_t, c_int64, c_int64, c_int64, c_void_p]
lib.whvi_layer_bwd_f32.restype = c_int
dev = torch.device("cuda:0")
x, dy = torch.randn(S, B, D, device=dev), torch.randn(S, B, D, device=dev)
g, s1, s2 = torch.randn(S, D, device=dev), torch.randn(D, device=dev), torch.randn(D, device=dev)
dx, dg, ds1, ds2 = torch.empty_like(x), torch.empty_like(g), torch.empty_like(s1), torch.empty_like(s2)
need = c_size_t(0)
lib.whvi_layer_bwd_workspace_bytes(S, B, D, ctypes.byref(need))
ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
def run():
    rc = lib.whvi_layer_bwd_f32(x.data_ptr(), B * D, dy.data_ptr(), g.data_ptr(), s1.data_ptr(), s2.data_ptr(), dx.data_ptr(),
                                dg.data_ptr(), ds1.data_ptr(), ds2.data_ptr(), None, ws.data_ptr(), ws.numel(), S, B, D, None)
    assert rc == 0, rc
for _ in range(3): run()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): run()
b.record(); b.synchronize()
ms = a.elapsed_time(b) / 10
print(f"{sys.argv[1]} D={D}: bwd {ms:.3f} ms  {12.0 * S * B * D / ms / 1e6:.0f} GB/s")
