import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; data=rows[hdr+1:]
ki=H.index('Kernel Name'); vi=H.index('Metric Value')
names=[(r[ki],float(r[vi].replace(',',''))/1e3) for r in data if len(r)>vi]
half=names[len(names)//2:]
agg=collections.OrderedDict()
for n,t in half:
    k=n[:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=t
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 8]: print(f'{v[1]:9.1f} us  x{v[0]:3d}  {k}')
print('total',sum(v[1] for v in agg.values()))
