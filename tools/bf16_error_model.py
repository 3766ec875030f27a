#!/usr/bin/env python
"""What error do bf16 OPERANDS alone produce?  A PyTorch model (CPU) of a synthetic cfg run twice: in fp32
(double-precision convolutions) and with only the convolution inputs and weights rounded to bf16 (fp32 accumulate,
fp32 epilogue, fp32 residual stream) - the arithmetic contract of the tcgen05 kernels.  Prints, per layer, the worst
per-image max-normalised error and the relative L2 error between the two: the floor any bf16 implementation sits
on, to compare with what tests/test_baseline_batches_gpu.py measures on the GPU against the reference.

    python tools/bf16_error_model.py resnet50 2 256
"""
import sys, numpy as np, torch, torch.nn.functional as F, struct
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[1]))
from sr_object_detection_b200 import synth
torch.set_num_threads(8)
name, B, side = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg=synth.CFGS[name](batch=B,w=side,h=side)
# parse cfg sections
secs=[]
for raw in cfg.splitlines():
    line="".join(raw.split())
    if not line or line[0] in "#;": continue
    if line.startswith("["): secs.append((line,{}))
    else:
        k,_,v=line.partition("="); secs[-1][1][k]=v
import tempfile, os
tmp=tempfile.mkdtemp(); synth.write_weights(tmp+"/w", cfg, seed=1234)
buf=np.fromfile(tmp+"/w",np.float32)[4:]
x0=torch.from_numpy(synth.images(B,3,side,side,seed=42))
def bf(t): return t.to(torch.bfloat16).to(torch.float32)
def run(emul):
    pos=0; outs=[]; x=x0; c=3
    for name_,o in secs[1:]:
        if name_=="[convolutional]":
            n=int(o["filters"]); k=int(o["size"]); s=int(o.get("stride",1)); pad=k//2 if int(o.get("pad",0)) else 0
            bn=int(o.get("batch_normalize",0)); cin=x.shape[1]
            nonlocal_buf=buf
            b=torch.from_numpy(buf[pos:pos+n].copy()); pos+=n
            if bn:
                sc=torch.from_numpy(buf[pos:pos+n].copy()); pos+=n
                mu=torch.from_numpy(buf[pos:pos+n].copy()); pos+=n
                var=torch.from_numpy(buf[pos:pos+n].copy()); pos+=n
            w=torch.from_numpy(buf[pos:pos+n*cin*k*k].copy()).reshape(n,cin,k,k); pos+=n*cin*k*k
            xi = bf(x) if emul else x
            wi = bf(w) if emul else w
            y=F.conv2d(xi.double() if not emul else xi, wi.double() if not emul else wi, None, s, pad).float()
            if bn:
                a=sc/(var.sqrt()+1e-6); y=y*a.view(1,-1,1,1)+(b-mu*a).view(1,-1,1,1)
            else: y=y+b.view(1,-1,1,1)
            if o.get("activation")=="leaky": y=torch.where(y>0,y,0.1*y)
            x=y
        elif name_=="[maxpool]":
            k=int(o["size"]); s=int(o["stride"]); x=F.max_pool2d(x,k,s)
        elif name_=="[shortcut]":
            f=outs[len(outs)+int(o["from"])]
            y=x.clone()
            if f.shape[2]!=x.shape[2]:
                st=f.shape[2]//x.shape[2]; f=f[:,:,::st,::st]
            cm=min(f.shape[1],x.shape[1]); y[:,:cm]+=f[:,:cm]
            if o.get("activation")=="leaky": y=torch.where(y>0,y,0.1*y)
            x=y
        else: break
        outs.append(x)
    return outs
ref=run(False); em=run(True)
for i,(a,b) in enumerate(zip(ref,em)):
    a=a.reshape(B,-1); b=b.reshape(B,-1)
    e=((a-b).abs().max(1).values/a.abs().max(1).values).max().item()
    l2=((a-b).double().pow(2).sum()/a.double().pow(2).sum()).sqrt().item()
    print(i, secs[i+1][0], 'max-norm %.2e L2 %.2e'%(e,l2))
