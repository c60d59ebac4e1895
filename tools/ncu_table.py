import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]
cur={}
for r in rows[h+1:]:
    d=dict(zip(hdr,r))
    k=(int(d['ID']),d['Kernel Name'][:48])
    cur.setdefault(k,{})[d['Metric Name']]=d['Metric Value']
for k,v in sorted(cur.items()):
    print(k[0], k[1], ' '.join(f"{m.split('.')[0].split('__')[-1]}={x}" for m,x in v.items()))
