"""Development aid: a few steps / launches over every kernel family (tensor-core and FFMA training, resident-set graph,
short and long trajectory generation) at small sizes."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "defensive-model-vae_b200")): sys.path.insert(0,p)
from dmvae import ConditionalTrajectoryVAE
from dmvae.train import FusedTrainer
torch.manual_seed(0)
for (T,L,B) in ((10,8,300),(10,32,260),(21,16,130)):
    m=ConditionalTrajectoryVAE(T,3,L).to("cuda"); tr=FusedTrainer(m, lr=1e-4)
    x=torch.randn(B,T,3,device="cuda").cumsum(1)
    for _ in range(2): tr.step(x)
    data=torch.randn(2*B,T,3,device="cuda").cumsum(1)
    gs=tr.capture(B, dataset=data)
    for _ in range(3): gs.replay()
    torch.cuda.synchronize(); print("train ok",T,L,B,[float(v) for v in tr.losses.cpu()][:2])
for (T,L,B) in ((10,8,300),(50,8,200),(400,64,130)):
    m=ConditionalTrajectoryVAE(T,3,L).to("cuda")
    o=m.generate(torch.tensor([[11.0,0.0]]), n=B, seed=1); o2=m.generate(torch.rand(B,2)*10, n=B, seed=2)
    torch.cuda.synchronize(); print("decode ok",T,L,B,float(o.abs().mean()))
# long trajectories on the tensor cores (chunked first / last layer), the metric scans, the MPC tracker, a lone sub-module
for (T,L,B) in ((50,8,300),(100,16,260)):
    m=ConditionalTrajectoryVAE(T,3,L).to("cuda"); tr=FusedTrainer(m, lr=1e-4)
    x=torch.randn(B,T,3,device="cuda").cumsum(1)
    for _ in range(2): tr.step(x)
    torch.cuda.synchronize(); print("train (long) ok",T,L,B,[float(v) for v in tr.losses.cpu()][:2])
from dmvae import validation as V
from dmvae.tracker import track_batch
import numpy as np
m=ConditionalTrajectoryVAE(10,3,8).to("cuda")
traj=m.generate(torch.tensor([[11.0,0.0]]), n=4096, seed=3)
v,(lo,hi)=V.waypoint_speeds(traj); print("metrics ok", lo, hi, int(V.trajectories_per_cell(traj,"vae_offset_sce4_cond").sum()))
t=np.arange(10)*0.7
way=np.repeat(np.stack([0.1*np.sin(t), 8*t-0.25*t*t, t],1)[None],64,0).astype(np.float32)
res=track_batch(way, np.repeat(np.array([[0,0,np.pi/2,0.1,8.0]]),64,0), 0.02, max_steps=20)
print("tracker ok", res.states[0,-1].tolist(), int(res.iterations[0]))
print("sub-module ok", tuple(m.encoder(torch.randn(5,10,3)).shape), tuple(m.fc_mu(torch.randn(5,256)).shape))
