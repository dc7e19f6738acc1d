"""Development aid: one scoring call at the README shape (optionally with the -DKCVAE_TAIL_TIMING build's
per-warp wait-cycle dump, KCVAE_TAIL_DBG=1).  usage: tools/tail_dbg.py [frames]"""
import sys, os, importlib, torch
sys.path.insert(0, "/root/repo")
from oracle import kcvae_oracle as O
pkg = importlib.import_module("trustedai-cl-vae-ad_b200")
cfg = O.readme_config(); B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
m = pkg.load_model_from_config(cfg, precision="bf16"); m.set_weights(O.glorot_init(cfg))
x = torch.rand(B,224,300,3,device="cuda")
for i in range(2): m.score(x)
torch.cuda.synchronize()
print("tc status", m.tc_status() if hasattr(m, "tc_status") else None)
os.environ["KCVAE_TAIL_DBG"]="tail"
m.score(x)
torch.cuda.synchronize()
