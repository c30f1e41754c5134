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
