"""Scratch: per-tile timeline of cluster 0 of the 2-CTA GEMM (library built with -DGOAT_TIMELINE)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vln_goat_b200 import ops, _lib
dev = "cuda"; dt = torch.bfloat16
M = 5120
x = torch.randn(M, 768, device=dev).to(dt); x32 = torch.randn(M, 768, device=dev)
W1 = (torch.randn(3072, 768, device=dev) * 0.02).to(dt); b1 = torch.zeros(3072, device=dev)
Wqkv = (torch.randn(2304, 768, device=dev) * 0.02).to(dt); b3 = torch.zeros(2304, device=dev)
Wo = (torch.randn(768, 768, device=dev) * 0.02).to(dt); b0 = torch.zeros(768, device=dev)
z = torch.empty(M, 3072, device=dev, dtype=dt)
cases = {
  "ffn1": lambda: ops.gemm(x, W1, bias=b1, act=ops.ACT_GELU, aux_out=z),
  "qkv": lambda: ops.gemm(x, Wqkv, bias=b3),
  "out": lambda: ops.gemm(x, Wo, bias=b0, res=x32, out_dtype=torch.float32),
}
lib = _lib.lib()
buf = (ctypes.c_longlong * 4096)()
names = ["mma_tile_start", "mma_tmem_empty_ok", "mma_first_full", "mma_tile_done", "epi_wait_start", "epi_acc_ready", "epi_tile_done"]
for name, fn in cases.items():
    for _ in range(3): fn()
    torch.cuda.synchronize()
    fn(); torch.cuda.synchronize()
    lib.goat_debug_timeline(buf)
    t0 = buf[7 * 64 + 2]
    print("== %s: kernel start 0, setup done %d, end %d (clocks)" % (name, buf[7*64+0]-t0, buf[7*64+1]-t0))
    for i in range(8):
        print("  chunk %d: ld_done %d staged %d stored %d" % (i, buf[8*64+i]-t0, buf[9*64+i]-t0, buf[10*64+i]-t0))
    for i in range(3):
        print("  tile %d: " % i + "  ".join("%s %d" % (names[s], buf[s*64+i]-t0) for s in range(7)))
