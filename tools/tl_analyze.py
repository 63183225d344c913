"""Summarise a GSP_PROF_TIMELINE dump (stderr lines 'TL dev stream name start_ms dur_ms') of a distributed factorization."""
import collections, sys
rows = []
for l in open(sys.argv[1]):
    if l.startswith("TL "):
        _, dev, st, name, t0, d = l.split()
        rows.append((int(dev), st, name, float(t0), float(d)))
devs = sorted(set(r[0] for r in rows))
end = max(r[3] + r[4] for r in rows)
print(f"{len(rows)} launches on {len(devs)} devices, span {end:.2f} ms")
for dv in devs:
    rs = [r for r in rows if r[0] == dv]
    sts = collections.OrderedDict()
    for r in sorted(rs, key=lambda r: r[3]):
        sts.setdefault(r[1], []).append(r)
    out = []
    for i, (st, lst) in enumerate(sts.items()):
        busy = sum(r[4] for r in lst)
        names = collections.Counter(r[2] for r in lst)
        out.append(f"s{i}: {busy:6.1f} ms busy, {len(lst)} launches ({', '.join(f'{k} {v}' for k, v in names.items())})")
    # union busy time over all streams
    ev = sorted((r[3], r[3] + r[4]) for r in rs)
    u, cur0, cur1 = 0.0, None, None
    for a, b in ev:
        if cur1 is None or a > cur1:
            if cur1 is not None:
                u += cur1 - cur0
            cur0, cur1 = a, b
        else:
            cur1 = max(cur1, b)
    u += cur1 - cur0
    print(f"dev {dv}: any-stream busy {u:6.1f} ms of {end:.1f}; " + " | ".join(out))
if len(sys.argv) > 2:
    dv = int(sys.argv[2]); lo = float(sys.argv[3]); hi = float(sys.argv[4])
    sts = {}
    for r in sorted(rows, key=lambda r: r[3]):
        if r[0] == dv:
            sts.setdefault(r[1], len(sts))
    for r in sorted(rows, key=lambda r: r[3]):
        if (dv < 0 or r[0] == dv) and lo <= r[3] <= hi:
            print(f"{r[3]:9.3f} {r[4]:7.3f} dev{r[0]} s{sts.get(r[1], '?')} {r[2]}")
