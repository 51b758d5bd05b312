import sys, os, copy, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visuelle2_multimodal_fusion_b200 import trunk
from visuelle2_multimodal_fusion_b200.models._base import resnet101_trunk
CL = torch.channels_last
torch.manual_seed(0)
torch.backends.cudnn.allow_tf32 = False
warnings.simplefilter("ignore")
truth = resnet101_trunk().cuda().train()
cnn = copy.deepcopy(truth).to(memory_format=CL)
ref = copy.deepcopy(truth).to(memory_format=CL)
x = torch.randn(8, 3, 299, 299, device="cuda")
def cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float(a @ b / (a.norm() * b.norm() + 1e-30))
mt, mc, mr = list(truth.children()), list(cnn.children()), list(ref.children())
with torch.no_grad():
    a = mt[3](mt[2](mt[1](mt[0](x))))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        b = mr[3](mr[2](mr[1](mr[0](x.contiguous(memory_format=CL)))))
        c = mc[3](trunk.bn_act(mc[0](x.contiguous(memory_format=CL)), mc[1], relu=True))
    print("stem   cos torch-bf16 %.6f fused %.6f" % (cos(b, a), cos(c, a)))
    for li in range(4, 8):
        for bi in range(len(mt[li])):
            a = mt[li][bi](a)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                b = mr[li][bi](b)
                c = trunk._bottleneck(mc[li][bi], c)
            print("layer%d.%d cos torch-bf16 %.6f fused %.6f  fused-vs-torchbf16 %.6f" % (li - 3, bi, cos(b, a), cos(c, a), cos(c, b)))
