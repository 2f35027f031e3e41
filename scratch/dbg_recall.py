import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import outline_rag_b200 as orx
from outline_rag_b200.sharded import ShardedIndex
from outline_rag_b200.synth import Synth
syn=Synth(1024)
n=20000
X=syn.table(n); Q,_=syn.queries(8,n)
a=ShardedIndex("bf16", n, 0); b=ShardedIndex("fp32", n, 0)
a.local.use_torch_stream(); b.local.use_torch_stream()
ids=np.zeros((n,2),np.uint64); ids[:,1]=np.arange(n)
xd=torch.from_numpy(X).cuda()
a.local.upsert(ids, xd); b.local.upsert(ids, xd)
print(len(a), len(b))
qd=torch.from_numpy(Q).cuda()
ga=a.search(qd,12); gb=b.search(qd,12)
print(ga[0][0,:,1], gb[0][0,:,1], ga[2], gb[2])
