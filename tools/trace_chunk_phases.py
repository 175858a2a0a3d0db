"""Average time between consecutive per-chunk trace points of the first epilogue warp (needs a -DDHG_TRACE_FINE build:
DHG_LIB_PATH=... DHG_NVCC_FLAGS=-DDHG_TRACE_FINE python -m dhg_b200.build; then DHG_TRACE=1 tools/gemm_one.py ... 2> log).  python tools/trace_chunk_phases.py log"""
import sys
ev=[]
for ln in open(sys.argv[1]):
    f=ln.split()
    if f and f[0]=='TR': ev.append((int(f[1]),int(f[2],16),int(f[3])))
ev.sort()
seq=[(t,c) for t,c,i in ev if 0x33<=c<=0x3b and c!=0x37]
names={0x33:'chunk start',0x34:'ld done',0x38:'before cp wait',0x39:'cp wait done',0x3a:'aux added',0x3b:'aux refill issued',0x35:'math done',0x36:'stored'}
acc={};cnt={}
for (t0,c0),(t1,c1) in zip(seq,seq[1:]):
    k=(c0,c1); acc[k]=acc.get(k,0)+t1-t0; cnt[k]=cnt.get(k,0)+1
for k in sorted(acc, key=lambda k:-cnt[k]):
    if cnt[k]>10: print(f"{names[k[0]]:20s} -> {names[k[1]]:20s} {acc[k]/cnt[k]:7.0f}  x{cnt[k]}")
