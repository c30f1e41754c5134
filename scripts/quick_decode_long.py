import os, sys, torch
ROOT="/root/repo"
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")): sys.path.insert(0,p)
from dmvae import ConditionalTrajectoryVAE
for T in (50,100,200,400):
    torch.manual_seed(0)
    m=ConditionalTrajectoryVAE(T,3,8).to("cuda")
    B=1<<20
    st=torch.tensor([[11.0,0.0]])
    m.generate(st,n=B,seed=1); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): m.generate(st,n=B,seed=1)
    e1.record(); torch.cuda.synchronize()
    dt=e0.elapsed_time(e1)/3*1e-3
    H=128; dec=(8+H)*H+2*H*H+3*T*H
    print(T, f"{B/dt:.4g} traj/s {B*2*dec/dt/1e12:.1f} TFLOP/s {B*T*12/dt/1e9:.0f} GB/s")
