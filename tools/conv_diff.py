"""Per-layer difference of two tools/conv_profile.py outputs."""
import sys
def load(f):
    d = {}
    for l in open(f).read().splitlines()[1:]:
        t = l.split()
        d[' '.join(t[6:])] = float(t[0])
    return d
a, b = load(sys.argv[1]), load(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 12
rows = sorted(((b.get(k, 0) - v, k, v, b.get(k, 0)) for k, v in a.items()))
for r in rows[:n] + rows[-n:]:
    print("%+.3f  %s  %.3f -> %.3f" % r)
print("total %.2f -> %.2f" % (sum(a.values()), sum(b.values())))
