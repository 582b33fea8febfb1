import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from graphnet_classifier_b200 import ops
from graphnet_classifier_b200.models.GNN import CombinedModel, GraphNet
from graphnet_classifier_b200.pipeline import GraphClassifierPipeline
r = 128
torch.manual_seed(0)
model = CombinedModel(GraphNet(num_local_features=3, space_dim=2, out_channels=1, n_blocks=3), num_nodes=r * r).cuda().eval()
pipe = GraphClassifierPipeline(model, resize_value=r)
img = torch.from_numpy(np.random.default_rng(0).integers(0, 256, (1, r, r, 3), dtype=np.uint8)).cuda()
for _ in range(3): pipe.infer(img)
ops.PROFILE = ops.KernelProfile(); pipe.infer(img); torch.cuda.synchronize()
for k, d in sorted(ops.PROFILE.summary().items(), key=lambda kv: -kv[1]["ms"]): print(f"{k:22s} {d['ms']:7.3f} ms x{d['calls']}")
