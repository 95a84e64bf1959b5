"""Executed instructions per event by source FUNCTION of one kernel (ncu SASS page x nvdisasm -g line info x the
source files' function starts).  Development tool.
  python tools/sass_by_function.py sass.csv lib.dis <kernel substring> <events> <source root that /root/repo maps to>"""
import csv,re,sys,collections,bisect
sass_csv,dis,kern,ev,srcroot=sys.argv[1],sys.argv[2],sys.argv[3],float(sys.argv[4]),sys.argv[5]
rows=list(csv.reader(open(sass_csv))); hdr=rows[1]; ix,ax=hdr.index("Instructions Executed"),hdr.index("Address")
cnt={}; base=None
for r in rows[2:]:
    if len(r)<=ix: continue
    a=int(r[ax],16); base=a if base is None else base; cnt[a-base]=int(r[ix] or 0)
line_of={}; inside=False; cur=None
for ln in open(dis):
    if ln.startswith(".text."): inside=kern in ln; continue
    if not inside: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)',ln)
    if m: cur=(m.group(1),int(m.group(2))); continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);",ln)
    if m: line_of[int(m.group(1),16)]=cur
funcs={}
def fn_table(path):
    if path in funcs: return funcs[path]
    starts=[];names=[]
    try:
        p=path.replace('/root/repo',srcroot)
        for i,l in enumerate(open(p),1):
            m=re.match(r"\s*(?:template\s*<[^>]*>\s*)?(?:static\s+)?(?:MSIM_HD|MSIM_NI|MSIM_FI|__device__|static MSIM_HD|inline)\s+.*?\b([A-Za-z_0-9]+)\s*\(",l)
            if m and not l.strip().startswith('//'): starts.append(i);names.append(m.group(1))
    except Exception as e: pass
    funcs[path]=(starts,names); return funcs[path]
agg=collections.Counter(); tot=0
for a,c in cnt.items():
    f,l=line_of.get(a,("?",0)); st,nm=fn_table(f)
    j=bisect.bisect_right(st,l)-1
    agg[(f.split('/')[-1], nm[j] if j>=0 else '?')]+=c; tot+=c
for k,v in agg.most_common(60): print(f"{k[0]:14s} {k[1]:24s} {v/ev:7.1f}/event {100*v/tot:5.1f}%")
