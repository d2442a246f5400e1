"""numpy emulation of the fp16 tensor-core field: which rounding points dominate the sigma error?"""
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests')
import numpy as np, torch, scenes
from geneface_b200 import synthetic
from oracle import field as OF, cpu_ops as ops
model, hp = synthetic.build_model(torso=False, bitfield='S', seed=0, device='cpu')
sd = synthetic.state_to_numpy(model)
xyz, d = scenes.field_samples(20000, seed=25, bound=1.0)
cf = torch.randn(64, generator=torch.Generator().manual_seed(3)).numpy().astype(np.float64)
fo = OF.FieldOracle(sd, bound=1.0)
h16 = lambda x: np.asarray(x, np.float32).astype(np.float16).astype(np.float64)
ident = lambda x: np.asarray(x, np.float64)
relu = lambda v: np.maximum(v, 0)
Wa = [sd[f'ambient_net.net.{i}.weight'].astype(np.float64) for i in range(3)]
Ws = [sd[f'sigma_net.net.{i}.weight'].astype(np.float64) for i in range(3)]
pos = OF.grid_encode(xyz, 1.0, sd['position_embedder.embeddings'], fo.pos_offsets, fo.pos_pls).astype(np.float64)
bias = Wa[0][:, 32:] @ cf

def run(rf=ident, ra=ident, rw=ident, rlast_a=None, rlast_w=None, ramb=None):
    """rf: feature rounding, ra: activation rounding, rw: weight rounding; rlast_*: override for the sigma-logit dot"""
    rlast_a = rlast_a or ra; rlast_w = rlast_w or rw; ramb = ramb or (rf, ra, rw)
    f_, a_, w_ = ramb
    h = relu(f_(pos) @ w_(Wa[0][:, :32]).T + bias)
    h = relu(a_(h) @ w_(Wa[1]).T)
    amb = np.tanh(a_(h) @ w_(Wa[2]).T).astype(np.float32)
    af = OF.grid_encode(amb, 1, sd['ambient_embedder.embeddings'], fo.amb_offsets, fo.amb_pls).astype(np.float64)
    h = relu(np.concatenate([rf(pos), rf(af)], 1) @ rw(Ws[0]).T)
    h = relu(ra(h) @ rw(Ws[1]).T)
    logit = rlast_a(h) @ rlast_w(Ws[2][0])
    return logit

ref = run()
def rep(name, **kw):
    e = np.abs(run(**kw) - ref)
    print(f"{name:55s} logit abs err: median {np.median(e):.2e} p99 {np.percentile(e,99):.2e} max {e.max():.2e}")
rep("all fp16 (current kernel)", rf=h16, ra=h16, rw=h16)
rep("all fp16, sigma logit dot in fp32", rf=h16, ra=h16, rw=h16, rlast_a=ident, rlast_w=ident)
rep("only features fp16", rf=h16)
rep("only activations fp16", ra=h16)
rep("only weights fp16", rw=h16)
rep("ambient branch exact, sigma branch fp16", rf=h16, ra=h16, rw=h16, ramb=(ident, ident, ident))
rep("ambient branch fp16 only", ramb=(h16, h16, h16))
rep("fp16 everything but weights exact", rf=h16, ra=h16)

print("---- per-layer toggles inside the ambient branch (sigma branch fp16) ----")
def run_amb(l0=(h16,h16), l1=(h16,h16), l2=(h16,h16)):
    h = relu(l0[0](pos) @ l0[1](Wa[0][:, :32]).T + bias)
    h = relu(l1[0](h) @ l1[1](Wa[1]).T)
    amb = np.tanh(l2[0](h) @ l2[1](Wa[2]).T).astype(np.float32)
    af = OF.grid_encode(amb, 1, sd['ambient_embedder.embeddings'], fo.amb_offsets, fo.amb_pls).astype(np.float64)
    h = relu(np.concatenate([h16(pos), h16(af)], 1) @ h16(Ws[0]).T)
    h = relu(h16(h) @ h16(Ws[1]).T)
    return h16(h) @ h16(Ws[2][0]), amb
ref_l, ref_amb = run_amb((ident,ident),(ident,ident),(ident,ident))
ex=(ident,ident)
def hl(x):  # hi+lo split fp16 (two-term), i.e. ~22 bit operand
    hi = h16(x); return hi + h16(np.asarray(x,np.float64)-hi)
sp=(hl,hl)
for name,kw in [("all ambient layers fp16",{}),("amb2 exact",dict(l2=ex)),("amb1+amb2 exact",dict(l1=ex,l2=ex)),("amb0 exact only",dict(l0=ex)),
                ("amb0+amb1 split(hi/lo), amb2 exact",dict(l0=sp,l1=sp,l2=ex)),("all split",dict(l0=sp,l1=sp,l2=sp)),
                ("amb1 split, amb2 exact, amb0 fp16",dict(l1=sp,l2=ex)),
                ("amb0: feat fp16 & W split; amb1 split; amb2 exact",dict(l0=(h16,hl),l1=sp,l2=ex))]:
    lg, amb = run_amb(**kw)
    e=np.abs(lg-ref_l); ea=np.abs(amb-ref_amb)
    print(f"{name:52s} logit err median {np.median(e):.2e} p99 {np.percentile(e,99):.2e} max {e.max():.2e} | amb_pos err max {ea.max():.2e}")
