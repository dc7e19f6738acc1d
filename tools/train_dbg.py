"""Development aid: train steps at the README shape with the -DKCVAE_TAIL_TIMING build; KCVAE_TAIL_DBG=<kernel tag>
(tail, out_dgrad, ...) prints that kernel's per-warp wait cycles on the last step.  usage: tools/train_dbg.py <tag> [frames]"""
import sys, os, importlib, torch
sys.path.insert(0, "/root/repo")
from oracle import kcvae_oracle as O
pkg = importlib.import_module("trustedai-cl-vae-ad_b200")
cfg = O.readme_config(); B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
m = pkg.load_model_from_config(cfg, precision="bf16"); m.set_weights(O.glorot_init(cfg)); m.compile(optimizer=pkg.Adam(1e-4))
x = torch.rand(B,224,300,3,device="cuda")
for i in range(3): m.train_step(x)
torch.cuda.synchronize()
os.environ["KCVAE_TAIL_DBG"] = sys.argv[1]
m.train_step(x)
torch.cuda.synchronize()
