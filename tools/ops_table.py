"""Group the per-launch timing table written by `bench.py --dump-ops` by conv shape."""
import collections
import json
import sys

ops = json.load(open(sys.argv[1]))
tot = sum(o['ms'] for o in ops)
print("total ms %.3f  launches %d" % (tot, len(ops)))
groups = collections.OrderedDict()
for o in ops:
    k = (o['kind'], o['cin'], o['cout'], o['ksize'], o['stride'], o['out_h'], o['out_w'])
    g = groups.setdefault(k, dict(n=0, ms=0, flops=0, bytes=0, ex=o))
    g['n'] += 1; g['ms'] += o['ms']; g['flops'] += o['flops']; g['bytes'] += o['bytes']
print(f"{'kind':8s} {'cin':>4s} {'cout':>4s} k s {'HxW':>7s} {'n':>3s} {'ms':>8s} {'%':>5s} {'TF/s':>7s} {'GB/s':>7s} {'us/launch':>9s} grid mb nt ck aS bS tiles")
for k, g in sorted(groups.items(), key=lambda kv: -kv[1]['ms']):
    e = g['ex']
    print(f"{k[0]:8s} {k[1]:4d} {k[2]:4d} {k[3]} {k[4]} {k[5]:3d}x{k[6]:<3d} {g['n']:3d} {g['ms']:8.3f} {100*g['ms']/tot:5.1f} "
          f"{g['flops']/g['ms']/1e9:7.1f} {g['bytes']/g['ms']/1e6:7.0f} {1e3*g['ms']/g['n']:9.1f} {e['grid']:4d} {e['mb']} "
          f"{e['nt']:3d} {e['ck']:2d} {e['a_stages']} {e['b_stages']} {e['tiles']}")
