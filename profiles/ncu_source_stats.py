import csv,sys,collections,re
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Address"); isrc=hdr.index("Source"); iex=hdr.index("Instructions Executed"); ismp=hdr.index("# Samples")
tot=sum(int(r[iex]) for r in data); cells=float(sys.argv[2])
print("total warp instr",tot,"per warp-cell",tot/cells)
c=collections.Counter(); s=collections.Counter()
for r in data:
    op=re.sub(r'^@!?U?P\d\s+','',r[isrc]).split()[0].split('.')[0]
    c[op]+=int(r[iex]); s[op]+=int(r[ismp])
ts=sum(s.values())
for op,n in c.most_common(28): print("%-10s %7.1f /warp-cell  samples %5.1f%%"%(op,n/cells,100*s[op]/ts))
# hottest instructions by samples
print("--- hottest by stall samples")
for r in sorted(data,key=lambda r:-int(r[ismp]))[:int(sys.argv[3]) if len(sys.argv)>3 else 25]:
    print(r[ia][-5:], "%5.2f%%"%(100*int(r[ismp])/ts), "exec/wc %.2f"%(int(r[iex])/cells), r[isrc][:70])
