import re,collections,sys
f=sys.argv[1]
rows=[tuple(map(int,re.findall(r'-?\d+',l))) for l in open(f) if l.startswith('trace')]
tot=collections.Counter(); cnt=collections.Counter()
for i,tag,dt in rows:
    tot[tag]+=dt; cnt[tag]+=1
T=sum(tot.values())
print(f,len(rows),'total cycles',T, 'us@1.965GHz', T/1965)
for t in sorted(tot): print(' tag',t,'n',cnt[t],'sum',tot[t],'avg',tot[t]//cnt[t], '%.1f%%'%(100*tot[t]/T))
print(' '.join('%d:%d'%(r[1],r[2]) for r in rows[200:260]))
