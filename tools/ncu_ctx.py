"""Like ncu_regions.py, but samples that land in inlined helpers (mbarrier waits, pack/unpack) are attributed to the last
tc4_gemm.cuh / given-file line seen before them in address order (the call site's neighbourhood = the warp role).
usage: ncu_ctx.py <mangled kernel> <src.csv> <segment> [file]"""
import re,csv,sys,subprocess,collections
kern=sys.argv[1]; srccsv=sys.argv[2]; which=int(sys.argv[3]); ctxfile=sys.argv[4] if len(sys.argv)>4 else 'tc4_gemm.cuh'
dis=subprocess.run(['nvdisasm','-g','-c','/tmp/cub/api.sm_100a.cubin'],capture_output=True,text=True).stdout.splitlines()
start=None
for i,l in enumerate(dis):
    if l.startswith('\t.section\t.text.'+kern): start=i
    elif start is not None and l.startswith('\t.section') and i>start: end=i; break
else: end=len(dis)
cur=None; ctx=None; line_of=[]
for l in dis[start:end]:
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2)))
        if cur[0]==ctxfile: ctx=cur[1]
        continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);',l)
    if m: line_of.append((int(m.group(1),16),ctx,cur,m.group(2)))
rows=list(csv.reader(open(srccsv)))
hdr=rows[1]; ia=hdr.index('Address'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
offmap={o:(c,cu,t) for o,c,cu,t in line_of}
starts=[i for i,r in enumerate(rows) if r and r[0]=='Kernel Name']
seg=rows[starts[which]+2:(starts[which+1] if which+1<len(starts) else len(rows))]
base=int(seg[0][ia],16)
inst=collections.Counter(); samp=collections.Counter(); tot=0; tots=0
for r in seg:
    off=int(r[ia],16)-base
    c=offmap.get(off,(None,None,None))[0]
    n=int(r[ie]); s=int(r[isamp])
    inst[c]+=n; samp[c]+=s; tot+=n; tots+=s
print('total inst',tot,'samples',tots)
for c in sorted(inst, key=lambda c:c or 0):
    if inst[c]/tot>0.004 or samp[c]/tots>0.004:
        print(f'{ctxfile}:{c}  inst {100*inst[c]/tot:5.1f}%  samples {100*samp[c]/tots:5.1f}%')
