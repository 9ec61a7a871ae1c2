"""Hottest SASS lines of an `ncu --page source --csv` export: python scripts/ncu_hot.py file_src.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[hi]; ix = {k: i for i, k in enumerate(h)}
body = [r for r in rows[hi + 1:] if len(r) >= len(h) - 2]
def f(r, k="Warp Stall Sampling (All Samples)"):
    try: return float(r[ix[k]])
    except (ValueError, IndexError): return 0.0
tot = sum(f(r) for r in body) or 1
print("total samples", tot, "instructions", len(body))
for i, r in enumerate(body):
    r.append(i)
top = sorted(body, key=f, reverse=True)[:n]
for r in top:
    st = sorted(((f(r, k), k[6:]) for k in h if k.startswith("stall_") and "Not Issued" not in k), reverse=True)[:2]
    print("%5d %6.2f%%  %-90s %s" % (r[-1], 100 * f(r) / tot, r[ix["Source"]][:90], " ".join("%s=%d" % (k, v) for v, k in st if v > 0)))
