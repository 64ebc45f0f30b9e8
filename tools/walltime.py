import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kmtricks_b200 import _lib, engine, synth
N=int(sys.argv[1]) if len(sys.argv)>1 else 20
MODE=sys.argv[2] if len(sys.argv)>2 else ''
if 'torch' in MODE:
    import torch; torch.cuda.set_device(0); torch.zeros(1,device='cuda')
R=1_000_000; Lr=150
cfg = engine.Config(kmer_size=31, nb_partitions=64, mode="hash:bf:bin", hard_min=2, bloom_size=200_000_000)
eng = engine.Engine(cfg, N); L=eng.lib; h=eng.h
sb = R*synth.record_bytes(Lr)
d_text=C.c_void_p(); L.kmx_dev_alloc(h, N*sb+64, C.byref(d_text))
for s in range(N): L.kmx_synth_fastq(h,1234,s,0,R,Lr,5_000_000,2e-3,2e-3,1,d_text.value+s*sb)
L.kmx_sync(h)
soft=np.full(N,1,dtype=np.uint32)
mp=_lib.KmxMergeParams(soft.ctypes.data_as(C.POINTER(C.c_uint32)),1,0,2,0); res=_lib.KmxMergeResult()
T={}
def tm(name, fn, *a):
    t0=time.perf_counter(); rc=fn(*a); T[name]=T.get(name,0)+time.perf_counter()-t0
    assert rc==0, (name, L.kmx_last_error(h))
if 'prof' in MODE: L.kmx_profile_enable(h,1)
for it in range(3):
    T.clear()
    t0=time.perf_counter()
    tm('reset', L.kmx_reset, h)
    for s in range(N):
        tm('begin', L.kmx_superk_begin, h)
        tm('push', L.kmx_superk_push_fastq, h, d_text.value+s*sb, sb, 1)
        tm('end', L.kmx_superk_end, h, None)
        tm('count', L.kmx_count_sample, h, s, 2)
    for p in range(64):
        tm('merge', L.kmx_merge_partition, h, p, C.byref(mp), C.byref(res))
    L.kmx_sync(h)
    print(it, 'total %.1f ms'%(1e3*(time.perf_counter()-t0)), {k: round(1e3*v,1) for k,v in T.items()})
