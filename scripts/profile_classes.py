"""One classification fit on device-resident synthetic data, for ncu captures of csrc/classes.cu.

    python scripts/profile_classes.py [rows] [features] [classes] [depth] [gini|p_at_k]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from np_decision_tree_b200.decision_tree import strategy, tree_builder  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
F = int(sys.argv[2]) if len(sys.argv) > 2 else 256
C = int(sys.argv[3]) if len(sys.argv) > 3 else 8
D = int(sys.argv[4]) if len(sys.argv) > 4 else 4
crit = strategy.greedy_classification_p_at_k if (len(sys.argv) > 5 and sys.argv[5] == "p_at_k") else strategy.greedy_classification
g = torch.Generator(device="cuda")
g.manual_seed(0)
X = torch.randn((N, F), dtype=torch.float64, device="cuda", generator=g)
labels = (X[:, :C] + 0.5 * torch.randn((N, C), dtype=torch.float64, device="cuda", generator=g)).argmax(dim=1).to(torch.int32)
t = tree_builder.build_classification_tree(X, labels, max_depth=D, max_feature=F, min_improvement=1e-7,
                                           split_method=crit, max_classes=C)
torch.cuda.synchronize()
print("nodes", t.max_node_ + 1)
