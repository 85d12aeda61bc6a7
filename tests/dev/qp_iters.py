import sys; sys.path.insert(0,'/root/repo')
import numpy as np
import hunter_bipedal_control_b200 as hb
from hunter_bipedal_control_b200 import scenarios as S
from oracle import hbo
import bench
B=256
x0,xr,sw,md,rbd=bench.workload(B)
ctx=hb.Context(max_batch=B)
xt,ut=ctx.mpc_cold_start(x0,md)
xt1,ut1,info=ctx.mpc_solve(x0,xr,sw,md,xt,ut)
al=0.2; xd=(1-al)*xt1[:,0]+al*xt1[:,1]; ud=(1-al)*ut1[:,0]+al*ut1[:,1]
H=np.zeros((B,38,38)); g=np.zeros((B,38)); A=np.zeros((B,60,38)); lb=np.full((B,60),-1e20); ub=np.full((B,60),1e20)
for i in range(B):
    Hi,gi,Ai,lbi,ubi=hbo.wbc_assemble(xd[i],ud[i],rbd[i],int(md[i,0]),False); m=Ai.shape[0]
    H[i]=Hi; g[i]=gi; A[i,:m]=Ai; lb[i,:m]=lbi; ub[i,:m]=ubi
x,st,it=ctx.wbc_qp(H,g,A,lb,ub)
print('generic QP iters mean %.2f min %d max %d'%(it.mean(),it.min(),it.max()), 'line search trials', info['n_trials'].mean(), 'alpha', info['alpha'].mean())
