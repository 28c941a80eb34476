"""Summarise an `ncu --page source --csv` dump of a warp-specialised kernel: per address range (role) the share of
executed instructions and of stall samples, and the hottest stall sites.  usage: ncu_roles.py dump.csv [lo:hi:name ...]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ia, isrc, iex, ismp = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
base, out = None, []
for r in rows[2:]:
    a = int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])
    if base is None:
        base = a
    out.append((a - base, r[isrc], int(r[iex]), int(r[ismp]), {hdr[i]: int(r[i] or 0) for i in stall_cols}))
key = ('UTCHMMA', 'UTMALDG', 'UTMASTG', 'LDTM', 'STTM', 'UTCBAR', 'BAR.SYNC', 'EXIT', 'UTMACMDFLUSH', 'UTMACCTL', 'DEPBAR')
if len(sys.argv) == 2:
    for o in out:
        if any(k in o[1] for k in key):
            print(hex(o[0]), o[1][:100].strip(), o[2], o[3])
    print('total inst', sum(o[2] for o in out), 'samples', sum(o[3] for o in out))
    sys.exit(0)
tot_i, tot_s = sum(o[2] for o in out), sum(o[3] for o in out)
for spec in sys.argv[2:]:
    lo, hi, name = spec.split(':')
    lo, hi = int(lo, 16), int(hi, 16)
    sel = [o for o in out if lo <= o[0] < hi]
    print(f"== {name}: inst {100 * sum(o[2] for o in sel) / tot_i:.1f}%  samples {100 * sum(o[3] for o in sel) / tot_s:.1f}%")
    for o in sorted(sel, key=lambda o: -o[3])[:10]:
        top = sorted(o[4].items(), key=lambda kv: -kv[1])[:2]
        print('  ', hex(o[0]), o[1][:64].strip(), o[2], o[3], top)
