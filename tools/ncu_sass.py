"""Per-instruction view of an .ncu-rep source page: executed count, stall samples, SASS.
usage: ncu_sass.py report.ncu-rep [min_exec]   (prints instructions whose executed count >= min_exec)"""
import csv, io, subprocess, sys
def load(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = None; data = []
    for r in rows:
        if r and r[0] == 'Address': hdr = r; continue
        if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
    return data
if __name__ == '__main__':
    d = load(sys.argv[1])
    mn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    tot_i = sum(int(x['Instructions Executed']) for x in d); tot_s = sum(int(x['# Samples']) for x in d)
    print(f"# {len(d)} instructions, {tot_i} executed, {tot_s} samples")
    for i, x in enumerate(d):
        ex = int(x['Instructions Executed'])
        if ex >= mn:
            print(f"{i:5d} {ex:10d} {float(x['Avg. Threads Executed'] or 0):5.1f} {int(x['# Samples']):6d}  {x['Source'].strip()}")
