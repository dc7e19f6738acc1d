import sys, importlib, torch
sys.path.insert(0, "/root/repo")
from oracle import kcvae_oracle as O
pkg = importlib.import_module("trustedai-cl-vae-ad_b200")
cfg = O.readme_config(); B = 32
m = pkg.load_model_from_config(cfg, precision="bf16"); m.set_weights(O.glorot_init(cfg)); m.compile(optimizer=pkg.Adam(1e-4))
xs = [torch.rand(B,224,300,3,device="cuda") for _ in range(6)]
for i in range(5): m.train_step(xs[i%6])
torch.cuda.synchronize()
e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K=20
e0.record()
for i in range(K): m.train_step(xs[i%6])
e1.record(); torch.cuda.synchronize()
print("step ms", e0.elapsed_time(e1)/K)
m.profile(True)
for i in range(K): m.train_step(xs[i%6])
rep = m.profile_report(); m.profile(False)
tot = sum(v[1] for v in rep.values())/K
print("sum of launcher times per step ms", tot, "launcher calls/step", sum(v[0] for v in rep.values())/K)
for k,v in sorted(rep.items(), key=lambda kv:-kv[1][1]): print(f"{k:40s} {v[0]//K:3d} {v[1]/K:8.4f}")
import time
t=time.perf_counter()
for i in range(K): m.train_step(xs[i%6])
print("host enqueue ms/step", (time.perf_counter()-t)/K*1e3); torch.cuda.synchronize()
