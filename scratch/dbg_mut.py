import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'oracle')); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np, torch
from world import World
from genetic_coverage_planning_b200 import synth
import test_gpu_variation as T
w=World.load_npz(os.path.join(ROOT,'tests/golden/world_big.npz'))
eng=T._engine(w, T._op(w, mutation_prob=1.0, muterpolate_prob=0.0))
par=T._parents(w, eng, 96, walk=120)
dna=torch.from_numpy(par.copy()).cuda(); eng.mutate(dna, seed=5, gen=1); out=dna.cpu().numpy()
edges=T._edges(w)
n=0
for p in range(96):
    for a in range(6):
        r0=par[p,a][par[p,a]>=0]; r1=out[p,a][out[p,a]>=0]
        for x in range(len(r1)-1):
            if (int(r1[x]),int(r1[x+1])) not in edges:
                n+=1
                if n<=3:
                    k=0
                    while k<min(len(r0),len(r1)) and r0[k]==r1[k]: k+=1
                    print('org',p,'agent',a,'bad at',x,'common prefix',k,'len0',len(r0),'len1',len(r1))
                    print(' r0',r0[max(0,x-6):x+8].tolist()); print(' r1',r1[max(0,x-6):x+8].tolist())
print('violations',n)
