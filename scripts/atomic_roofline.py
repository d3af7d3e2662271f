import sys; sys.path.insert(0,'/root/repo')
from radiosity_b200 import api
s=api.Scene(0.014)
for k,N in ((64,512),(1,512)):
    ctx=api.context_for_scene(s,N,k,select_mode=api.SELECT_TOPK if k>1 else 0)
    for pat in (0,1,2):
        print('k',k,'pattern',pat,[round(ctx.bench_atomics(pat, 1<<27),1) for _ in range(3)],'G RED.MIN.64/s')
    ctx.close()
