"""print the clock64 stamps of the probe build of attention_tc.cu (LS_SO_NAME=_C_probe.so)"""
import os, sys, ctypes as C, torch
os.environ.setdefault("LS_SO_NAME", "_C_probe.so")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from latentsync_b200 import _lib as L
dev = "cuda:0"
batch, heads, d, S = 32, 8, 40, 1024
Cc = heads * d
qkv = (torch.randn(batch * S, 3 * Cc, device=dev) * 1.5).half()
out = torch.empty(batch * S, Cc, dtype=torch.float16, device=dev)
for _ in range(3):
    L.attention(qkv[:, :Cc], qkv[:, Cc:2 * Cc], qkv[:, 2 * Cc:], out, 3 * Cc, 3 * Cc, 3 * Cc, Cc, batch, heads, d, S, S)
torch.cuda.synchronize()
n = 16 * 64 * 16
buf = (C.c_longlong * n)()
assert L.lib().ls_atc_probe_read(buf, n) == 0
names = ["c:S_issue", "c:PV_issue", "c:pv_done", "s:wait_S", "s:got_S", "s:ld_done", "s:exp_done", "s:p_arrive",
         "c:top", "c:s_full", "c:ldK", "c:k_full", "c:S_issued", "c:v_full", "c:PV_issued", "c:ldV"]
order = [3, 4, 5, 6, 7, 8, 9, 10, 0, 12, 11, 1, 14]
for cta in (0, 8):
    base = buf[(cta * 64 + 0) * 16 + 3]
    print(f"cta slot {cta}: stamps relative to softmax tile-0 start; columns: " + " ".join(names[k] for k in order))
    for j in range(16):
        v = [buf[(cta * 64 + j) * 16 + k] for k in order]
        print(f"  tile {j:2d}: " + " ".join(f"{(x - base) if x else -1:7d}" for x in v))
