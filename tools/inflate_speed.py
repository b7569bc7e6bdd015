"""Throughput of the native single-stream gzip decoder (kmb_gzstream_*) against zlib on synthetic FASTQ, one core.
Usage: python tools/inflate_speed.py"""
import ctypes as C
import gzip
import os
import sys
import time
import zlib

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_mapper_b200 import _lib
L=_lib.lib()
rng=np.random.default_rng(7)
n=400000
seq=rng.choice(np.frombuffer(b"ACGT",dtype=np.uint8),size=(n,150))
q=rng.choice(np.frombuffer(b"FFFFFFFF:,#",dtype=np.uint8),size=(n,150))
hdr=np.frombuffer(b"@r0000000/1\n",dtype=np.uint8)
rec=np.concatenate([np.broadcast_to(hdr,(n,len(hdr))), seq, np.full((n,1),10,np.uint8), np.broadcast_to(np.frombuffer(b"+\n",dtype=np.uint8),(n,2)), q, np.full((n,1),10,np.uint8)],axis=1)
data=rec.tobytes()
for level in (1,6):
    blob=gzip.compress(data,level)
    g=np.frombuffer(blob,dtype=np.uint8)
    out=np.empty(len(data)+(1<<20),dtype=np.uint8)
    best_z=1e9; best_k=1e9
    for rep in range(3):
        t=time.perf_counter(); z=zlib.decompress(blob,31); best_z=min(best_z,time.perf_counter()-t)
        for thr in (8,):
            h=C.c_void_p(); L.kmb_gzstream_open(g.ctypes.data,len(blob),thr,C.byref(h))
            prod=C.c_uint64(); fin=C.c_int()
            t=time.perf_counter()
            rc=L.kmb_gzstream_read(h,out.ctypes.data+65536,out.shape[0]-65536,0,C.byref(prod),C.byref(fin))
            dt=time.perf_counter()-t
            assert rc==0 and fin.value==1 and prod.value==len(data)
            L.kmb_gzstream_close(h); best_k=min(best_k,dt)
    assert out[65536:65536+len(data)].tobytes()==data
    print("level",level,"ratio %.1f"%(len(data)/len(blob)),"zlib %.3f GB/s  kmb(decode on 1 core, crc on the pool) %.3f GB/s"%(len(data)/best_z/1e9,len(data)/best_k/1e9))
