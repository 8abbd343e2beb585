import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "cs-vit_b200"))
import torch
from cs_vit import ops
hw, c = 56, 128
x = torch.randn(256 * hw * hw, c, device="cuda"); g = torch.ones(c, device="cuda"); b = torch.zeros(c, device="cuda")
ops.layernorm(x, g, b, 1e-5, out_dtype=torch.bfloat16); torch.cuda.synchronize()
