import sys, hashlib
sys.path.insert(0,'/root/repo')
import numpy as np
from tests import refdata
from tests.test_gpu_pipeline import build_blocks
from vorbis_aotuv_lancer_b200 import ops
cap = refdata.capture("q4_44k_2ch", 6.0)
s = ops.Setup("q4_44k_2ch")
blocks, pcm = build_blocks(cap, s.channels)
ref = cap["packets"]
res=[]
for trial in range(4):
    st = ops.Stream(s); bt = ops.Batch(s, len(blocks))
    half = len(blocks)//2 if trial<3 else len(blocks)
    bt.run([st], (ops.Block*half)(*blocks[:half]), pcm); pk = bt.packets()
    if half < len(blocks):
        bt.run([st], (ops.Block*(len(blocks)-half))(*blocks[half:]), pcm); pk += bt.packets()
    res.append([p[0] for p in pk])
    diff=[i for i,(a,r) in enumerate(zip(pk,ref)) if a[0]!=r[0]]
    print(trial, sum(len(p[0]) for p in pk), hashlib.md5(b''.join(p[0] for p in pk)).hexdigest(), 'ndiff vs ref', len(diff), diff[:20])
    st.close(); bt.close()
for t in range(1,4):
    d=[i for i,(a,b) in enumerate(zip(res[0],res[t])) if a!=b]
    print('trial',t,'vs 0: differing blocks', d[:30])
