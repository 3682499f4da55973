"""Compare two builds of the library kernel by kernel (SASS text): which kernels of build A are not byte-identical in build B, and which still
differ once kernel-parameter (constant bank 0) offsets are masked.  Used to show that an off-by-default addition left the kernels of a
GPU-verified build unchanged.   python tools/sass_compare.py old.so new.so"""
import subprocess,hashlib,re,sys
def funcs(path):
    out=subprocess.run(['cuobjdump','-sass',path],capture_output=True,text=True).stdout
    d={}; name=None; body=[]
    for line in out.splitlines():
        m=re.match(r'\s*Function : (\S+)',line)
        if m:
            if name: d[name]='\n'.join(body)
            name=m.group(1); body=[]; continue
        m=re.match(r'\s*/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\*',line)
        if m: body.append(m.group(1))
    if name: d[name]='\n'.join(body)
    return d
norm=lambda t: re.sub(r'c\[0x0\]\[0x[0-9a-f]+\]','c[0x0][*]',t)
a=funcs(sys.argv[1]); b=funcs(sys.argv[2])
hb_exact={hashlib.md5(v.encode()).hexdigest() for v in b.values()}
hb_norm={hashlib.md5(norm(v).encode()).hexdigest() for v in b.values()}
ex=[n for n,v in a.items() if hashlib.md5(v.encode()).hexdigest() not in hb_exact]
nm=[n for n in ex if hashlib.md5(norm(a[n]).encode()).hexdigest() not in hb_norm]
print(len(a),'functions before,',len(b),'after; not byte-identical:',len(ex),'; still different after masking kernel-parameter offsets:',len(nm))
for n in ex: print('  param-offset only' if n not in nm else '  DIFFERENT', n[:110])
