import sys, time, ctypes as CT
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pastml_b200 import synth, _lib
from pastml_b200.engine import pack_params
rng=np.random.default_rng(4000)
K,C=20,8
flat=synth.yule_tree(1_000_000,rng)
masks,has,pis,sfs=synth.simulate_tip_masks(flat,K,C,rng)
params=np.stack([pack_params(0,K,sfs[c],0.,1.,pis[c]) for c in range(C)])
def pinned(shape,dt):
    t=torch.empty(shape,dtype=dt,pin_memory=True); return t,t.numpy()
keep=[]
def P(shape,dt):
    t,a=pinned(shape,dt); keep.append(t); return a
h_parent=P(flat.n_nodes,torch.int32); h_dist=P(flat.n_nodes,torch.float64); h_masks=P((C,flat.n_nodes),torch.int64); h_has=P((C,flat.n_nodes),torch.uint8)
h_parent[:]=flat.parent; h_dist[:]=flat.dist; h_masks.view(np.uint64)[:]=masks; h_has[:]=has
h_probs=P((flat.n_nodes,C,K),torch.float64); h_map=P((flat.n_nodes,C),torch.int64); h_mppa=P((flat.n_nodes,C),torch.int64)
lib=_lib.load(); vp=lambda a: a.ctypes.data_as(CT.c_void_p)
h_lh=np.empty(C); h_hist=np.empty((C,K+1),dtype=np.int64)
for it in range(4):
    t0=time.perf_counter(); handle=CT.c_void_p()
    rc=lib.pml_problem_create(flat.n_nodes,vp(h_parent),vp(h_dist),C,K,0,vp(h_masks),vp(h_has),1,0,CT.byref(handle)); assert rc==0
    t1=time.perf_counter()
    rc=lib.pml_marginal(handle,vp(params),0,vp(h_lh),vp(h_probs),None,vp(h_map),vp(h_mppa),vp(h_hist)); assert rc==0
    t2=time.perf_counter()
    lib.pml_problem_destroy(handle); t3=time.perf_counter()
    print('create %.1f ms  marginal %.1f ms  destroy %.1f ms'%((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3))
