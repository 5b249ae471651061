"""Critical path of one ADMM iteration of the one-per-SM CTA kernel from ncu's per-instruction stall samples
(ncu --set full --import-source on ... tools/prof_solo.py; ncu -i X.ncu-rep --page source --csv > X.csv).
python tools/critical_path.py X.csv CYCLES_PER_ITERATION  -> per role: the loop split at its named barriers, samples and
barrier-stall samples per segment, converted to cycles with the role's share of the iteration."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) == len(h)]
cyc = float(sys.argv[2])
S, E, B = ix['# Samples'], ix['Instructions Executed'], ix['stall_barrier']
names = {'0x1': 'first exchange (kBarAll)', '0x3': 'y (kBarY)', '0x4': 'obstacle terms (kBarT)', '0x5': 'burst command', '0x6': 'hand-over to assistants (kBarH1)',
         '0x7': 'slack part rs (kBarRS)', '0x8': 'assistant level (kBarA)', '0x0': 'CTA'}
cnt = collections.Counter(r[E] for r in data)
# hot loops: runs of consecutive rows with the same large execution count
hot = [r[E].isdigit() and int(r[E]) > 20000 for r in data]
loops = []
i = 0
while i < len(data):
    if hot[i]:
        j = i; last = i
        while j < len(data) and (hot[j] or j - last <= 3):
            if hot[j]: last = j
            j += 1
        if last + 1 - i > 60: loops.append((i, last + 1, max(int(data[t][E]) for t in range(i, last + 1) if hot[t])))
        i = last + 1
    else: i += 1
for (a, b, e) in loops:
    tot = sum(int(data[i][S] or 0) for i in range(a, b))
    print(f"loop rows {a}..{b} ({b-a} instructions, executed {e} times each): {tot} samples = one iteration of {cyc:.0f} cycles")
    seg_start = a; label = 'loop top'
    def flush(lo, hi, label):
        s_ = sum(int(data[i][S] or 0) for i in range(lo, hi)); w = sum(int(data[i][B] or 0) for i in range(lo, hi))
        print(f"   {label:58s} {hi-lo:4d} instr  {s_:6d} samples ({cyc*s_/tot:6.0f} cycles), of which waiting at a barrier {w:6d} ({cyc*w/tot:6.0f} cycles)")
    for i in range(a, b):
        ins = data[i][1]
        if 'BAR.SYNC' in ins or 'BAR.ARV' in ins:
            bid = ins.split()[-2].rstrip(',') if ins.strip().endswith(';') else ins.split()[-2].rstrip(',')
            toks = ins.replace(',', ' ').split()
            bid = next((t for t in toks if t.startswith('0x')), '?')
            flush(seg_start, i + 1, label)
            seg_start = i + 1; label = ('after arrive: ' if 'ARV' in ins else 'after sync: ') + names.get(bid, bid)
    flush(seg_start, b, label)
