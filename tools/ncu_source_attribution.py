"""Attribute the per-instruction counters of an `ncu --set full --import-source on` capture to the inlined
__device__ functions they were compiled from.

    cuobjdump -xelf all rust_pathtracer_b200/libptb200.so && nvdisasm -g -c ptb_api.sm_100a.cubin > /tmp/lib.sass
    ncu -i X.ncu-rep --page source --csv > X_src.csv
    python tools/ncu_source_attribution.py X_src.csv _ZN3ptb18k_render_wavefrontILb0ELb0E [lines]

Joins the SASS offsets of the capture with nvdisasm's `//## File ..., line N` markers (needs -lineinfo, and the SAME
build as the capture) and prints, per function: share of warp-level instructions, average active lanes, share of
stall samples, static instruction count."""
import csv,re,sys,collections
csvf=sys.argv[1]; kern=sys.argv[2]
rows=list(csv.reader(open(csvf)))
hdr=rows[1]; data=rows[2:]
ia=hdr.index("Address"); ii=hdr.index("Instructions Executed"); it=hdr.index("Thread Instructions Executed"); isamp=hdr.index("# Samples")
base=int(data[0][ia],16)
per_off={}
for r in data:
    off=int(r[ia],16)-base
    per_off[off]=(int(r[ii]),int(r[it]),int(r[isamp]))
lines=open('/tmp/lib.sass').read().splitlines()
start=[i for i,l in enumerate(lines) if l.startswith('.text.'+kern)][0]
end=[i for i,l in enumerate(lines) if l.startswith('\t.section') and i>start][0]
cur=None; agg=collections.defaultdict(lambda:[0,0,0,0])
srcs={}
def func_of(f,line):
    if f not in srcs:
        try: srcs[f]=open('/root/repo/rust_pathtracer_b200/csrc/'+f).read().splitlines()
        except Exception: srcs[f]=[]
    src=srcs[f]
    for i in range(min(line,len(src))-1,-1,-1):
        m=re.match(r'^(template <[^>]*>\s*)?(PTB_DEV|PTB_HD|__device__ inline|__global__)\s.*?\b(\w+)\(',src[i])
        if m: return m.group(3)
    return f
bylines=collections.defaultdict(lambda:[0,0,0])
for line in lines[start:end]:
    m=re.search(r'//## File "([^"]+)", line (\d+)',line)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,6})\*/',line)
    if m:
        off=int(m.group(1),16)
        if off in per_off and cur:
            e,t,s=per_off[off]
            a=agg[func_of(*cur)]; a[0]+=e; a[1]+=t; a[2]+=s; a[3]+=1
            b=bylines[cur]; b[0]+=e; b[1]+=t; b[2]+=s
tot=sum(a[0] for a in agg.values()); tots=sum(a[2] for a in agg.values())
print(f"{'function':28s} {'warp-inst%':>10s} {'lanes':>6s} {'samples%':>8s} {'static':>6s}")
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:40]:
    print(f"{k:28s} {a[0]/tot*100:10.2f} {a[1]/max(a[0],1):6.1f} {a[2]/max(tots,1)*100:8.2f} {a[3]:6d}")
if len(sys.argv)>3:
    print("top lines")
    for k,b in sorted(bylines.items(), key=lambda kv:-kv[1][2])[:30]:
        print(k, b[0], round(b[1]/max(b[0],1),1), b[2])
