import csv, collections, re, sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit'); mi=hdr.index('Metric Name')
agg=collections.OrderedDict()
for r in rows[1:]:
    name=re.sub(r'\(.*','',r[ki]); name=re.sub(r'.*::','',name)
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1000
    elif r[ui]=='ms': v*=1000
    agg.setdefault((name,r[mi]),[]).append(v)
for (k,mname),v in agg.items():
    print('%-34s %-28s n=%4d  mean %10.2f  min %10.2f  max %10.2f  total %12.1f'%(k[:34],mname[:28],len(v),sum(v)/len(v),min(v),max(v),sum(v)))
