"""Runs one ResNet50 batch with TF2B_MMA_DEBUG=1 so conv_mma prints its per-role cycle breakdown."""
import os, sys
os.environ["TF2B_MMA_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tf2_b200.network import NetWork, Runner
net, q, model = bench.build_model()
nw = NetWork(net, 0)
nw.InitFromCodes(model, q, max_images=256)
x = torch.randint(-128, 128, (256, 3, 224, 224), dtype=torch.int8).cuda()
r = Runner(nw)
for i in range(2):
    print(f"--- pass {i}", file=sys.stderr)
    r.run_device(x, raw224=True)
torch.cuda.synchronize()
