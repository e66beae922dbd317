import sys, numpy as np
sys.path.insert(0,'.')
from starch3_b200 import synth
from oracle import oracle as O
cfg=int(sys.argv[1]); lines=int(sys.argv[2]); KB=int(sys.argv[3]) if len(sys.argv)>3 else 40
bed=synth.bed(cfg,lines).tobytes()
tfb,chs,_=O.transform(bed)
c=max(chs,key=lambda c:c['tf_len'])
s=tfb[c['tf_off']:c['tf_off']+c['tf_len']]
blk=np.frombuffer(s[:899981],dtype=np.uint8)
n=len(blk)
syms,inv=np.unique(blk,return_inverse=True)
a=len(syms); 
k=1; pw=a
while pw*a<=(1<<KB): pw*=a; k+=1
f=min((1<<KB)//pw,a)
print('n',n,'A',a,'k',k,'f',f)
ext=np.concatenate([inv,inv[:k+1]]).astype(np.uint64)
key=np.zeros(n,dtype=np.uint64)
for j in range(k): key=key*np.uint64(a)+ext[j:j+n]
key=key*np.uint64(f)+(ext[k:k+n]*np.uint64(f)//np.uint64(a))
for bits in ([KB, KB-10, KB-20] ):
    kk=key>>np.uint64(KB-bits)
    u,c=np.unique(kk,return_counts=True)
    print('bits',bits,'groups',len(u),'singletons',int((c==1).sum()),'entries in nonsingle',int(c[c>1].sum()),'max',c.max(),'sum sq (nonsingle)',int((c[c>1].astype(np.int64)**2).sum()))
    hist=np.bincount(np.minimum(np.ceil(np.log2(c)).astype(int),12),weights=c)
    print('   entries by log2 size',hist.astype(int).tolist())
    hist2=np.bincount(np.minimum(np.ceil(np.log2(c)).astype(int),12),weights=c.astype(np.float64)**2)
    print('   sumsq by log2 size',hist2.astype(np.int64).tolist())
# per-row loop lengths of the warp finisher
order=np.argsort(key,kind='stable'); ks=key[order]
head=np.ones(n,bool); head[1:]=ks[1:]!=ks[:-1]
hidx=np.flatnonzero(head); size=np.diff(np.append(hidx,n))
rows=hidx//32
mx=np.zeros(n//32+1,dtype=np.int64)
np.maximum.at(mx,rows,np.where(size>1,size,0))
print('rows',len(mx),'mean max group size per row',mx.mean(),'rows with nonsingle',(mx>0).mean(), 'mean size of nonsingle entries', (size[size>1]**2).sum()/size[size>1].sum())
