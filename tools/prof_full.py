"""Kernel-time breakdown of one compress+decompress step of the imagenet64.yaml model (development aid)."""
import os, sys, random
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from flic_b200 import flows
layer = dict(name="DenseLayer", act="ReLU")
block = dict(name="DenseBlock", growth_channel=512, depth=12, layer=layer)
cfg = dict(name="IDFlows", nflows=8, nbits=8, nsplit=3, H=64, W=64, C=3,
           couple=dict(name="AdditiveCouple", split=0.75, nn=block, round=dict(name="Round", nbits=8)),
           extenddim=dict(name="ExtendDim", scale=2), prior=dict(name="Prior", round=dict(name="Round", nbits=8), nn=block),
           distribution=dict(name="DLogistic"), round=dict(name="Round", nbits=8))
torch.manual_seed(0); random.seed(0)
model = flows.build_model(cfg); flows.perturb_heads(model, 0.02); model = model.cuda().eval()
img = torch.randint(0, 256, (256, 3, 64, 64), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234)).cuda()
def step():
    return model.decompress(model.compress(img, codec_batch=64, check=False), check=False)
assert torch.equal(step(), img)
step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
