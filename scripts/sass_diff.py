"""Per-kernel SASS comparison of two builds of libgopfcuda.so (instruction text, encodings, line-info comments
and anonymous-namespace hashes ignored).  Used to show that a refactor, or a compile-time switch such as
-DGOPF_KNOISE, leaves the kernels of a path untouched before any GPU time is spent on it.

    python scripts/sass_diff.py old/libgopfcuda.so new/libgopfcuda.so
"""
import re, collections, subprocess, sys
def split(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d={}; cur=None
    for line in out.splitlines():
        line=re.sub(r'/\* 0x[0-9a-f]* \*/','',line)
        line=re.sub(r'_GLOBAL__N__[0-9a-f]*_','_GLOBAL__N__X_',line)
        line=' '.join(line.split())
        if not line or line.startswith('//##'): continue
        m=re.search(r'Function : (\S+)', line)
        if m: cur=m.group(1); d[cur]=[]
        elif cur and re.match(r'/\*[0-9a-f]{4,}\*/', line): d[cur].append(line)
    return d
a=split(sys.argv[1]); b=split(sys.argv[2])
diff=[k for k in a if k in b and a[k]!=b[k]]
print(len(a), len(b), "differing:", len(diff), "only in one:", len(set(a)^set(b)))
for k in diff[:10]: print("  ", k[:110])
