import sys; sys.path.insert(0,'/root/repo')
import torch
from concepthash_b200 import synth, hashing
ev = hashing.get_evaluator()
for (nq,ndb,nbit,C) in [(2000,1_000_000,128,101),(2000,1_000_000,64,101),(2000,1_000_000,128,1000)]:
    d,dl,q,ql,_ = synth.make_random_case(nq,ndb,nbit,C,p=0.3,seed=0,device='cuda')
    m = ev.evaluate(d,dl,q,ql,[1000],0.0,[],False)
    print(nq,ndb,nbit,C,m[0],ev.stats)
