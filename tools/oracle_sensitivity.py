import sys, re, torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parent.parent))
from oracle import rspnet_oracle as oracle
from rspnet_b200.models import get_model_class
def cos(a,b): return float((a.flatten().double()@b.flatten().double())/(a.norm().double()*b.norm().double()+1e-30))
for arch,size in (("resnet18",64),("s3dg",128)):
    torch.manual_seed(0)
    net = get_model_class(arch=arch)(num_classes=1)
    sd = {"enc."+k:v.clone() for k,v in net.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4,3,8,size,size,generator=g)
    res=[]
    for trial,(emu,eps) in enumerate([(True,0.0),(True,1e-6),(False,0.0)]):
        oracle.EMULATE_BF16 = emu
        names=[k for k in oracle.param_names(sd,"enc.")]
        leaves={k:sd[k].clone().requires_grad_(True) for k in names}
        sdl=dict(sd); sdl.update(leaves)
        xx = x*(1+eps*torch.randn(x.shape, generator=torch.Generator().manual_seed(9)))
        feat=oracle.FEATURES[arch](oracle._r(xx), sdl, "enc.", True)
        if trial==0: R=torch.randn(feat.shape,generator=g)
        used=[k for k in names if not any(t in k for t in (".fc.",".linear."))]
        gr=torch.autograd.grad((feat*R).sum(),[leaves[k] for k in used],allow_unused=True)
        res.append((feat.detach(),dict(zip(used,gr))))
    oracle.EMULATE_BF16=False
    k0=[k for k in used if res[0][1][k] is not None][0]
    print(arch,"first param",k0)
    print("  emu vs emu(+1e-6 input noise): feat rel",((res[0][0]-res[1][0]).abs().max()/res[0][0].abs().max()).item(),"grad cos",cos(res[0][1][k0],res[1][1][k0]))
    print("  emu vs fp32:                   feat rel",((res[0][0]-res[2][0]).abs().max()/res[2][0].abs().max()).item(),"grad cos",cos(res[0][1][k0],res[2][1][k0]))
