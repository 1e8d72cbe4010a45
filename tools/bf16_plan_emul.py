"""CPU study (kernel emulation, tests/emul_backend.py) of bf16 operand plans for the tcgen05 GEMMs against the reference's
ResNet-101 goldens: activations as a sum of 1-3 bf16 terms, W+ (positive twin and dgrads) rounded to bf16, the signed forward
weights as 1-3 bf16 terms.  kind::f16 MMAs run at twice the TF32 rate on half-size shared-memory tiles; this only answers
whether the NUMERICS would hold.  Prints max-abs / max-abs over max(ref) per map."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import emul_backend as EB
from emul_backend import EmulBackend, im2col_nhwc, relu
from xfr_b200.packing import unpack_dual_cols
from helpers import L101, L1111, golden, golden_inputs, rel_err
from xfr_b200 import synth
from xfr_b200.engine import StResnetEngine
def bf16(w): return w.to(torch.bfloat16).to(torch.float32)
def bfsplit(x, terms):
    out=torch.zeros_like(x); r=x
    for _ in range(terms):
        h=bf16(r); out=out+h; r=r-h
    return out
W_=EB._w
class BF(EmulBackend):
    a_terms=2; wf_terms=2
    def _dgrad(self, y, Bd, R, signed=False):
        if signed: return super()._dgrad(y,Bd,R,signed)
        return im2col_nhwc(bfsplit(y,self.a_terms), R, R, R // 2) @ bf16(W_(Bd)).t()
    def conv_dual(self, inp, L, o, xr, act, res=None, relu_act=True):
        A = im2col_nhwc(bfsplit(inp,self.a_terms), L.R, L.S, L.R // 2)
        Wfull=W_(L.Bf)
        t, _ = unpack_dual_cols(A @ bfsplit(Wfull,self.wf_terms).t() + L.bias, L.tn)
        _, p = unpack_dual_cols(A @ bf16(Wfull).t() + L.bias, L.tn)
        o.view(-1, L.cout).copy_(t); xr.view(-1, L.cout).copy_(relu(p))
        a = t * L.bn[0] + L.bn[1]
        if res is not None:
            rc = res.shape[-1]; a[:, :rc] += res.reshape(-1, rc)
        act.view(-1, L.cout).copy_(relu(a) if relu_act else a)
def run(tag, a_terms, wf_terms, layers=L101):
    G=golden(layers); x,W2,_=golden_inputs(G)
    P1=torch.zeros(2,2); P1[:,0]=1
    be=BF(impl_name='tf32x3'); be.a_terms=a_terms; be.wf_terms=wf_terms
    eng=StResnetEngine(synth.stresnet_state_dict(0,layers,2), be, layers)
    s=eng.ebp(x,P1,W2).clone().numpy(); c=eng.contrastive(x,W2).clone().numpy(); t=eng.contrastive(x,W2,percentile=20).clone().numpy()
    for i,p in enumerate(('smooth','noise')):
        print(tag, p, 'ebp %.1e/%.1e | contrastive %.1e/%.1e | trunc %.1e/%.1e  (max-abs/rel)'%(np.abs(s[i]-G['ebp_awp_%s'%p]).max(), rel_err(s[i],G['ebp_awp_%s'%p]), np.abs(c[i]-G['cebp_awp_%s'%p]).max(), rel_err(c[i],G['cebp_awp_%s'%p]), np.abs(t[i]-G['tcebp20_awp_%s'%p]).max(), rel_err(t[i],G['tcebp20_awp_%s'%p])))
if __name__=='__main__':
    run('act bf16x3, fwdW bf16x3, W+ bf16', 3, 3)
    run('act bf16x2, fwdW bf16x2, W+ bf16', 2, 2)
    run('act bf16x2, fwdW bf16x1, W+ bf16', 2, 1)
    run('act bf16x1, fwdW bf16x1, W+ bf16', 1, 1)
