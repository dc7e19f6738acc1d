import sys, importlib, torch
sys.path.insert(0, "/root/repo")
from oracle import kcvae_oracle as O
pkg = importlib.import_module("trustedai-cl-vae-ad_b200")
cfg = O.readme_config(); B = 128
m = pkg.load_model_from_config(cfg, precision="bf16"); m.set_weights(O.glorot_init(cfg))
xs = [torch.rand(B,224,300,3,device="cuda") for _ in range(3)]
for i in range(3): m.score(xs[i%3])
torch.cuda.synchronize()
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K=10
e0.record()
for i in range(K): m.score(xs[i%3])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)/K
print("score ms per call", ms, "frames/s", B/ms*1e3)
m.profile(True)
for i in range(K): m.score(xs[i%3])
rep = m.profile_report(); m.profile(False)
print("sum ms", sum(v[1] for v in rep.values())/K)
for k,v in sorted(rep.items(), key=lambda kv:-kv[1][1]): print(f"{k:40s} {v[0]//K:3d} {v[1]/K:8.4f}")
