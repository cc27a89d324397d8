"""Compare two `cuobjdump -sass` dumps function by function: which kernels are byte-identical, changed, removed, new.
Used to show that adding opt-in kernel variants to a translation unit left every already-validated kernel untouched
(round 1: all 104 device functions of the GPU-validated commit are identical in the final build).

    cuobjdump -sass old/tc_conv.o > a.sass; cuobjdump -sass new/tc_conv.o > b.sass; python tools/sass_diff.py a.sass b.sass
"""
import re, sys
def funcs(path):
    out = {}
    cur = None
    for line in open(path):
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1); out[cur] = []
            continue
        if cur is not None:
            # drop addresses? keep instruction text + encoding
            out[cur].append(line.rstrip())
    return out
a = funcs(sys.argv[1]); b = funcs(sys.argv[2])
same = [k for k in a if k in b and a[k] == b[k]]
diff = [k for k in a if k in b and a[k] != b[k]]
gone = [k for k in a if k not in b]
new = [k for k in b if k not in a]
print("unchanged %d, changed %d, removed %d, new %d" % (len(same), len(diff), len(gone), len(new)))
for k in diff: print("CHANGED", k)
for k in gone: print("REMOVED", k)
for k in new: print("NEW", k)
