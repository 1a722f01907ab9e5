"""Summarises an ncu report's source page for one kernel: where the warp samples sit (mbarrier wait loops by barrier
offset, named-barrier stalls, the hottest instructions).  usage: ncu_stalls.py report.ncu-rep [bar_base_hex]"""
import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
S = lambda r: int(r[ix['# Samples']] or 0)
tot = sum(S(r) for r in data)
print("kernel:", rows[0][1][:100]); print("total samples", tot)
waits = collections.Counter()
for i, r in enumerate(data):
  if 'TRYWAIT' in r[ix['Source']]:
    m = re.search(r'\[(.*?)\]', r[ix['Source']])
    key = m.group(1) if m else '?'
    n = S(r)
    for j in range(i + 1, min(i + 3, len(data))):
      if 'BRA' in data[j][ix['Source']]:
        n += S(data[j]); break
    waits[key] += n
print("mbarrier wait loops: %d samples (%.1f%%)" % (sum(waits.values()), 100 * sum(waits.values()) / tot))
for k, v in waits.most_common(12): print("   %-28s %6d  %.1f%%" % (k, v, 100 * v / tot))
byexec = collections.Counter()
for r in data: byexec[r[ix['Instructions Executed']]] += S(r)
print("samples by execution count (phase fingerprint):")
for k, v in byexec.most_common(10): print("   execs %-10s %6d  %.1f%%" % (k, v, 100 * v / tot))
print("hottest instructions:")
for r in sorted(data, key=lambda r: -S(r))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
  print("   %6d %9s lsb %5s ssb %4s bar %4s  %s" % (S(r), r[ix['Instructions Executed']], r[ix['stall_long_sb']], r[ix['stall_short_sb']], r[ix['stall_barrier']], r[ix['Source']].strip()[:85]))
