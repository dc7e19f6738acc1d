"""Development aid: a few README-config train steps (and one scoring call) for ncu.  usage: tools/prof_step.py [frames] [cfg5]"""
import sys, os, importlib, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import kcvae_oracle as O
pkg = importlib.import_module("trustedai-cl-vae-ad_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
scaled = len(sys.argv) > 2 and sys.argv[2] == "cfg5"
cfg = O.scaled_config() if scaled else O.readme_config()
m = pkg.load_model_from_config(cfg)
if not scaled: m.set_weights(O.glorot_init(cfg))
m.compile(optimizer=pkg.Adam(1e-4))
H, W, C = cfg["data"]["image_size"]
x = torch.rand(B, H, W, C, device="cuda")
for i in range(3): m.train_step(x)
torch.cuda.synchronize()
m.score(x)
torch.cuda.synchronize()
print("ok", m.tc_status())
