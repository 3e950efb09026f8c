# summarise per-CTA phase stamps written by tools/stamps_mega.py (dev tool)
import sys, numpy as np
t=np.fromfile(sys.argv[1],dtype=np.int64).reshape(-1,128,4)[:148, :121]
t=t-t[:,0,0].min()
print('step total us', (t[:,120,3].max())/1e3)
def stats(x): return 'min %5.0f med %5.0f max %5.0f'%(x.min(),np.median(x),x.max())
tot=0
for p in range(4):
    phs=[l*4+p for l in range(5,25)]
    pro=t[:,phs,1]-t[:,phs,0]; gem=t[:,phs,2]-t[:,phs,1]; epi=t[:,phs,3]-t[:,phs,2]
    g=t[:,phs,3].max(0)- t[:,[x-1 for x in phs],3].max(0)
    tot+=g.mean()
    print('phase',p,'prologue(ns):',stats(pro),'| gemv:',stats(gem),'| epi:',stats(epi),'| global phase %.0f'%g.mean(), '| gemv-start spread %.0f'%((t[:,phs,1].max(0)-t[:,phs,1].min(0)).mean()))
print('per layer ns',tot)
l=12
for c in [0,1,74,147]:
    print(c, [(t[c,l*4+p]-t[:,l*4,0].min()).tolist() for p in range(4)])
