import csv,re,collections,sys
rows=list(csv.reader(open(sys.argv[1],errors='ignore')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value'); mu=h.index('Metric Unit')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    t=float(r[mv].replace(',',''))/1000
    name=r[kn]
    name=re.sub(r'at::native::','',name)
    m=re.search(r'(RowwiseMoments|CUDAFunctor_add|GroupNormKernel|CatArray|upsample|direct_copy|silu_kernel|avg_pool|ComputeFused|MulFunctor|layer_norm|LayerNorm|gelu|GeluCUDA|softmax|cos_kernel|sin_kernel|gemm|cutlass|sgemm|ampere|sm90|sm100|cudnn|implicit)',name)
    if m: name='torch:'+m.group(1)
    else: name=re.sub(r'\(.*','',name); name=re.sub(r'<.*','',name)
    agg[name[:60]][0]+=1; agg[name[:60]][1]+=t
tot=sum(v[1] for v in agg.values()); print('total us',round(tot,1))
for k,v in sorted(agg.items(),key=lambda kv:-kv[1][1])[:28]: print(f"{v[1]:9.1f} us {v[0]:5d}  {100*v[1]/tot:5.1f}%  {k}")
