"""What random 256-byte row access costs on this HBM: a gather of random rows (sml_gather_pairs: 2 x 256 B random reads +
512 B sequential write per id) and torch's index_select / index_copy on 20 M-row tables, against the streaming copy peak."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sml_b200 import ops
dev = torch.device("cuda:0")
n, m = 20_000_000, 2_000_000
a = torch.randn(n, 64, device=dev); b = torch.randn(n, 64, device=dev)
ids = torch.randint(0, n, (m,), device=dev)
uniq = torch.randperm(n, device=dev)[:m].contiguous()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {}
ms = timed(lambda: ops.gather_pairs(a, b, ids)); res["gather_pairs_gbs"] = m * 1024 / ms / 1e6
ms = timed(lambda: torch.index_select(a, 0, ids)); res["index_select_gbs"] = m * 512 / ms / 1e6
src = torch.randn(m, 64, device=dev)
ms = timed(lambda: a.index_copy_(0, uniq, src)); res["index_copy_gbs"] = m * 512 / ms / 1e6
c = torch.empty_like(a)
ms = timed(lambda: c.copy_(a)); res["stream_copy_gbs"] = 2 * n * 256 / ms / 1e6
print(json.dumps(res))
