import sys, os, torch
sys.path.insert(0, "/root/repo")
import hdiff_b200.ops as hops
ops = hops.get()
dev, bf = torch.device("cuda"), torch.bfloat16
N,H,W,C0,Cout = 4,128,128,128,128
x0 = torch.randn(N,H,W,C0,device=dev).to(bf)
w = (torch.randn(Cout,9,C0,device=dev)/30).to(bf)
bias = torch.randn(Cout,device=dev)
out = torch.empty(N,H,W,Cout,device=dev,dtype=bf)
for i in range(3):
    try:
        ops.conv(x0,None,1,w,bias,None,None,out,1,N,H,W,3)
        torch.cuda.synchronize()
        print("call", i, "ok", float(out.float().abs().mean()))
    except Exception as e:
        print("call", i, "ERR", str(e)[-120:])
