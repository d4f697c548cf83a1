import sys; sys.path.insert(0,'/root/repo')
import torch
from sml_b200 import ops
dev=torch.device('cuda:0')
for n_items in (122816, 5_000_000, 20_000_000):
    it=torch.randn(n_items,64,device=dev); ut=torch.randn(65536,64,device=dev)
    users=torch.randint(0,65536,(16384,),device=dev); pos=torch.randint(0,n_items,(16384,),device=dev)
    ipk=ops.pack_rows(it)
    for name,fn in (("rank",lambda: ops.fullcat_ranks(ut,it,users,pos,items_packed=ipk,n_items=n_items)),("topk20",lambda: ops.fullcat_topk(ut,it,users,20,items_packed=ipk,n_items=n_items)),("topk64",lambda: ops.fullcat_topk(ut,it,users,64,items_packed=ipk,n_items=n_items))):
        fn(); torch.cuda.synchronize()
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)
        print(n_items,name,"%.2f ms  %.1f TFLOP/s fp32-equiv"%(ms, 2*64*16384*n_items/ms/1e9), flush=True)
    del it,ipk; torch.cuda.empty_cache()
