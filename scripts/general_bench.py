import json, torch, sys
sys.path.insert(0, "/root/repo")
import bench
from cross_attention_renderer_b200 import _lib
ctx = {"dev": torch.device("cuda", 0)}
torch.cuda.set_device(0)
_lib.load()
for prec in ("fp32", "fp32_simt"):
    print(prec, json.dumps(bench.general_branches(ctx, 256, 64, 3, 2, precision=prec)), flush=True)
