import ctypes as C, numpy as np, time, torch, os
from gams_b200 import _lib, api, synth
L=_lib.load(); ctx=api.Context([0]); dev=torch.device('cuda',0)
lens=np.full(6200, 500000); n_q=100_000_000
d_st,d_sp,iv_off=synth.rg_index_torch(lens,16128,seed=7,device=dev)
q_seg,q_s,q_e=synth.queries_torch(lens,n_q,seed=11,grouped=False,device=dev)
def pinned(n): 
    p=L.gams_cuda_host_alloc(n); return p
ptrs=[]
for t in (d_st,d_sp,q_seg,q_s,q_e):
    pp=pinned(t.numel()*4); h=np.ctypeslib.as_array(C.cast(pp,C.POINTER(C.c_int32)),shape=(t.numel(),)); torch.from_numpy(h).copy_(t); ptrs.append(pp)
p_out=pinned(n_q*4); torch.cuda.synchronize()
ivp=iv_off.ctypes.data_as(C.c_void_p)
for i in range(3):
    t=time.perf_counter()
    api.check(L.gams_cuda_count_overlaps(ctx.handle, iv_off.size-1, ptrs[0],ptrs[1],ivp,n_q,ptrs[2],ptrs[3],ptrs[4],p_out))
    print('call', i, (time.perf_counter()-t)*1e3, 'ms', flush=True)
