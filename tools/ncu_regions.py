"""Per-role stall-sample split of a warp-specialised kernel capture: regions are delimited by marker opcodes."""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
b = src.split('"Kernel Name"')[1]
rr = list(csv.reader(io.StringIO('"Kernel Name"' + b)))
h = rr[1]; I = {x: i for i, x in enumerate(h)}
data = [r for r in rr[2:] if len(r) == len(h)]
marks = []
for i, r in enumerate(data):
    s = r[I['Source']]
    m = re.search(r'(LDGSTS|UTCHMMA|LDTM|BAR\.SYNC|UTCBAR|STG|SYNCS\.PHASECHK|SYNCS\.ARRIVE|DEPBAR|EXIT)', s)
    if m and int(r[I['Instructions Executed']]) > 0:
        marks.append((i, m.group(1), int(r[I['# Samples']]), int(r[I['Instructions Executed']])))
prev = 0
tot = sum(int(r[I['# Samples']]) for r in data)
print('total samples', tot)
for (i, name, smp, ex) in marks:
    seg = sum(int(r[I['# Samples']]) for r in data[prev:i + 1])
    print(f'{prev:5d}-{i:5d} {name:16s} exec {ex:9d} seg_samples {seg:5d}')
    prev = i + 1
print(f'{prev:5d}-end seg_samples', sum(int(r[I['# Samples']]) for r in data[prev:]))
