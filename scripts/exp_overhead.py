"""Where the time of one search step goes beyond the pair kernel (N = 1): wall clock around each piece, synchronised."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import vid_dup_finder_lib_b200 as vdf
from tests import synth
from vid_dup_finder_lib_b200 import _ffi, dist as vdist

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dev = torch.device("cuda", 0)
ctx = _ffi.Context(0)
H, _ = synth.planted_hashes(n)
dur = np.full(n, 600, np.uint32)
d_hash = torch.from_numpy(H.view(np.int64)).to(dev)
d_dur = torch.from_numpy(dur.view(np.int32)).to(dev)
torch.cuda.synchronize()
tbl = ctx.table_create_device(d_hash.data_ptr(), d_dur.data_ptr(), n, keepalive=(d_hash, d_dur))
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def wall(fn, k=10):
    fn(); torch.cuda.synchronize()
    t = []
    for _ in range(k):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
    return round(float(np.median(t)), 3)

res = {}
def keys_only():
    res["keys"] = vdist.search_self_keys(ctx, tbl, None, 350, device=dev)
def group_only():
    k = res["keys"]; res["g"] = ctx.group_greedy_device(n, k.data_ptr(), k.numel())
def both():
    keys_only(); torch.cuda.current_stream().synchronize(); group_only()
def one_call():
    res["g1"] = tbl.search_self_groups(350)
out = {"flush": wall(lambda: flush.fill_(1))}
ctx.kernel_time(0, reset=True)
out["keys_only"] = wall(keys_only)
ms, cnt = ctx.kernel_time(0, reset=True)
out["kernel"] = round(ms / cnt, 3)
out["group_only"] = wall(group_only)
out["both"] = wall(both)
out["one_call"] = wall(one_call)
out["edges"] = int(res["keys"].numel())
print(out)
