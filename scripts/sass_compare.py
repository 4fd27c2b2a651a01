"""Compares the SASS of every kernel of two builds of libbabyai_b200.so (cuobjdump -sass dumps), instruction text only.
Used when a change must not touch the code of kernels that were validated / profiled on the GPU: new behaviour goes into
new template instantiations (k_gen<true>, k_rollout<1, true>, k_step8<*, true>) and the existing ones must come out
IDENTICAL.

    cuobjdump -sass old.so > a.txt; cuobjdump -sass new.so > b.txt; python scripts/sass_compare.py a.txt b.txt
"""
import re, sys
def funcs(path):
    out={}; cur=None
    for line in open(path):
        m=re.search(r'Function : (\S+)', line)
        if m: cur=m.group(1); out[cur]=[]; continue
        if cur is None: continue
        m=re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\*', line)
        if m: out[cur].append(m.group(1))
    return out
a=funcs(sys.argv[1]); b=funcs(sys.argv[2])
def norm(n): return n.replace('ELb0EE','EE')
bmap={norm(k):k for k in b if 'Lb1' not in k or 'k_gen' in k}
for k in a:
    kb=bmap.get(k) or bmap.get(norm(k))
    if kb is None: print('missing', k); continue
    same = a[k]==b[kb]
    nd = sum(1 for x,y in zip(a[k],b[kb]) if x!=y) + abs(len(a[k])-len(b[kb]))
    print('%-75s %6d/%6d instrs  %s  (%d differing lines)' % (k[:75], len(a[k]), len(b[kb]), 'IDENTICAL' if same else 'DIFFERENT', nd))
