"""CPU study (kernel emulation, tests/emul_backend.py): how many mantissa bits do the tensors SAVED by the forward sweep (o =
conv + bias, xr = relu(conv+ + b'), read back by every hook chain of the backward) need?  They are shared by the mate and the
non-mate gradient rows, so their rounding largely cancels in the contrastive map.  ResNet-101 goldens; prints max-abs /
max-abs over max(ref) per map."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import emul_backend as EB
from emul_backend import EmulBackend
from helpers import L101, golden, golden_inputs, rel_err
from xfr_b200 import synth
from xfr_b200.engine import StResnetEngine
def bf16(w): return w.to(torch.bfloat16).to(torch.float32)
def f16ish(w, bits):   # keep `bits` mantissa bits (round to nearest)
    i=w.contiguous().view(torch.int32); sh=23-bits
    return ((i + (1<<(sh-1))) & ~((1<<sh)-1)).view(torch.float32)
class SB(EmulBackend):
    o_bits=None; xr_bits=None
    def conv_dual(self, inp, L, o, xr, act, res=None, relu_act=True):
        super().conv_dual(inp, L, o, xr, act, res, relu_act)
        if self.o_bits: o.copy_(f16ish(o, self.o_bits))
        if self.xr_bits: xr.copy_(f16ish(xr, self.xr_bits))
G=golden(L101); x,W2,_=golden_inputs(G)
P1=torch.zeros(2,2); P1[:,0]=1
for tag,ob,xb in (('xr 7 bits (bf16)',None,7),('xr 10 bits',None,10),('o,xr 7 bits',7,7),('o,xr 10 bits (fp16-like)',10,10),('o 15 bits, xr 7 bits',15,7)):
    be=SB(impl_name='tf32x3'); be.o_bits=ob; be.xr_bits=xb
    eng=StResnetEngine(synth.stresnet_state_dict(0,L101,2), be, L101)
    s=eng.ebp(x,P1,W2).clone().numpy(); c=eng.contrastive(x,W2).clone().numpy()
    for i,p in enumerate(('smooth','noise')):
        print('%-28s %-6s ebp %.1e/%.1e | contrastive %.1e/%.1e (max-abs/rel)'%(tag,p,np.abs(s[i]-G['ebp_awp_%s'%p]).max(), rel_err(s[i],G['ebp_awp_%s'%p]), np.abs(c[i]-G['cebp_awp_%s'%p]).max(), rel_err(c[i],G['cebp_awp_%s'%p])))
