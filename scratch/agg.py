import csv,re,collections,sys
def agg(path):
    rows=list(csv.reader(open(path,errors='ignore')))
    hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
    h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
    a=collections.defaultdict(lambda:[0,0.0])
    for r in rows[hi+1:]:
        if len(r)<=mv: continue
        t=float(r[mv].replace(',',''))
        if r[mu]=='ns': t/=1000
        name=r[kn]
        name=re.sub(r'at::native::','',name)
        m=re.search(r'(RowwiseMoments|CUDAFunctor_add|GroupNormKernel|CatArray|upsample|direct_copy|silu_kernel|avg_pool|ComputeFused|MulFunctor|cos_kernel|sin_kernel)',name)
        if m: name='torch:'+m.group(1)
        else:
            name=re.sub(r'\(.*','',name); 
        a[name[:70]][0]+=1; a[name[:70]][1]+=t
    return a
A=agg(sys.argv[1]); B=agg(sys.argv[2]) if len(sys.argv)>2 else None
tot=sum(v[1] for v in A.values()); print('total us',round(tot,1), 'prev', round(sum(v[1] for v in B.values()),1) if B else '')
for k,v in sorted(A.items(),key=lambda kv:-kv[1][1])[:26]:
    pv=B.get(k,[0,0.0]) if B else [0,0]
    print(f"{v[1]:9.1f} us {v[0]:5d}  {100*v[1]/tot:5.1f}%   prev {pv[1]:9.1f} {pv[0]:5d}  {k}")
