"""Per-SASS-instruction stall samples of an ncu source page (ncu -i X.ncu-rep --page source --csv > X.csv).
python tools/ncu_hot.py X.csv [first_row last_row]  -> top instructions, or the listing of a row range with samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
S = ix['# Samples']; IE = ix['Instructions Executed']
stall_cols = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
tot = sum(int(r[S] or 0) for r in data)
def line(i):
    r = data[i]
    st = {n[6:]: int(r[ix[n]] or 0) for n in stall_cols}
    main = sorted(((k, v) for k, v in st.items() if v > 0), key=lambda kv: -kv[1])[:3]
    return f"{i:6d} {r[0][-6:]} {r[1][:64]:64s} smp {int(r[S] or 0):6d} exe {r[IE]:>9s} " + " ".join(f"{k}:{v}" for k, v in main)
if len(sys.argv) > 3:
    a, b = int(sys.argv[2]), int(sys.argv[3])
    sub = sum(int(data[i][S] or 0) for i in range(a, b))
    print(f"rows {a}..{b}: {sub} samples of {tot} ({100*sub/tot:.1f}%)")
    for i in range(a, b): print(line(i))
else:
    print('total samples', tot, 'instruction rows', len(data))
    top = sorted(range(len(data)), key=lambda i: -int(data[i][S] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 60]
    for i in sorted(top): print(line(i))
